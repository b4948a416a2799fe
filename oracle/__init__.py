"""CPU oracle for the DeformContact message-passing hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``deformcontact_b200/`` imports this
package; it is imported by ``tests/``, by ``__graft_entry__.smoke()`` (as the
checker) and by ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
(as the thing timed on host cores).  It never sits on the product path.

What it restates
----------------
The reference (``/root/reference``) is 100 % Python and delegates all of the
hot-path arithmetic to third-party libraries that are neither vendored nor
installable here (no network, no wheel):

* ``torch_geometric`` pinned ``pyg=2.5.2`` (``environment.yml:76``):
  ``TAGConv`` / ``GCNConv`` / ``GATConv`` (``models/model.py:2,39``),
  ``gcn_norm``, ``MessagePassing.propagate``, ``Batch.from_data_list``
  (``train.py:36-38``), ``Data`` (``utils/graph_utils.py:20``);
* ``torch_cluster`` (no pin at all): ``knn_graph`` / ``radius_graph``
  (``utils/pointcloud_utils.py:10,12``);
* Open3D 0.18 ``create_sphere`` (``loaders/common.py:26``).

Each function below cites the reference call site it stands in for and
restates the published algorithm of the pinned third-party version in plain
PyTorch on the CPU (fp32 by default, fp64 on request as accuracy arbiter).

Parity status
-------------
**PARITY UNPINNED for the third-party arithmetic** (TAGConv/GCNConv/GATConv,
knn/radius graph, Batch): the reference has no tests, golden vectors or
fixtures, and the real libraries cannot be imported.
The oracle is instead cross-checked against (tests/test_oracle.py):
  (1) an independent dense fp64 formulation  sum_k A_hat^k X W_k^T + b,
  (2) scipy.sparse CSR matmul,
  (3) scipy.spatial.cKDTree for kNN / radius,
  (4) property tests (permutation equivariance, linearity, degenerate graphs),
  (5) a real ``torch_geometric`` if one is ever importable (skipped otherwise).
**Pinned against the reference itself** where the reference code is
importable in the build container: ``utils/pos_encoding.py:to_log_freq``,
``models/losses.py:GradientConsistencyLoss``, ``loaders/collate.py:collate_fn``
and the *wiring* of ``models/model.py:GraphNet`` (run with the oracle convs
injected for the missing ``torch_geometric.nn``).  Those outputs are committed
as fixtures under ``tests/golden/`` by ``tests/golden/make_golden.py``.
"""
from .data import Data, Batch  # noqa: F401
from .convs import (gcn_norm, propagate, TAGConv, GCNConv, GATConv,  # noqa: F401
                    MPNNLayer)
from .graphs import (mesh_to_graph, knn_graph, radius_graph, construct_graph,  # noqa: F401
                     to_log_freq, uv_sphere, grid_mesh, canonical_sort)
from .model import (GraphNet, MultiHeadAttention, GradientConsistencyLoss,  # noqa: F401
                    load_model, train_step_loss, EVERYDAY)
