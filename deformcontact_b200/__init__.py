"""deformcontact_b200 — B200-native (sm_100a) message-passing hot path of DeformContact.

Importing the package loads ``libdcb200.so`` (built in-tree by ``__graft_entry__.build()``);
there is no CPU or eager fallback — a missing library raises.
"""
from . import _abi

_abi.lib()  # fail loudly at import if the native library is absent

from . import ops  # noqa: E402
from .data import Data, Batch, collate_fn  # noqa: E402,F401
from .layers import TAGConv, GCNConv, GATConv, MPNNLayer, layer_stack  # noqa: E402,F401
from .graph import mesh_to_graph, knn_graph, radius_graph, construct_graph, to_log_freq  # noqa: E402,F401
from .assemble import (batch_from_data_list, mesh_batch, collider_batch, collider_batch_device, graph_batch,  # noqa: E402,F401
                       graph_batch_packed, Staging)
from .model import (GraphNet, MultiHeadAttention, GradientConsistencyLoss, load_model,  # noqa: E402,F401
                    train_step_loss, fused_losses, EVERYDAY)

from .step import CapturedTrainStep  # noqa: E402,F401

__version__ = "0.2.0"
