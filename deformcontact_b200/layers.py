"""Drop-in message-passing layers: ``TAGConv`` / ``GCNConv`` / ``GATConv`` (+ ``MPNNLayer``).

Same constructor signature ``(in_channels, out_channels)``, call ``layer(x, edge_index)`` and
state-dict keys as torch_geometric 2.5.2, so they slot into the reference's
``GraphNet.__init__`` / ``forward`` (models/model.py:39-50, 69-78) and load its checkpoints
(train.py:125, eval.py:36,89).  Forward and backward run entirely in libdcb200 (hand-written
sm_100a CUDA reached through the C ABI); every layer is ONE registered custom op with a registered
autograd formula (``torch.ops.dcb200.{tag_conv, gcn_conv, gat_conv, mpnn_layer}``, torch_ops.py) — the
modules below only hold the parameters and call the op.

``edge_index`` may be the usual ``int64 [2, E]`` tensor (the CSR pair is built once per tensor
and cached) or a prebuilt ``ops.GraphCSR`` (adopted into that cache under its own ``edge_index``).
"""
import math

import os

import torch
import torch.nn as nn

from . import ops
from . import torch_ops  # noqa: F401  (registers torch.ops.dcb200.*)


def _kaiming_uniform_linear_(w):
    bound = 1.0 / math.sqrt(w.size(-1)) if w.size(-1) > 0 else 0.0
    with torch.no_grad():
        w.uniform_(-bound, bound)


def _glorot_(w):
    a = math.sqrt(6.0 / (w.size(-2) + w.size(-1)))
    with torch.no_grad():
        w.uniform_(-a, a)


class _Lin(nn.Module):
    """Holds ``weight [out, in]`` under PyG's ``lins.k.weight`` / ``lin.weight`` key."""

    def __init__(self, i, o, init="kaiming"):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(o, i))
        (_glorot_ if init == "glorot" else _kaiming_uniform_linear_)(self.weight)


def _edge_tensor(edge_index):
    """The tensor the custom ops take: ``edge_index`` itself, or — for a prebuilt ``ops.GraphCSR`` — the tensor it was built
    from, after adopting the structure into the cache the ops look it up in."""
    if isinstance(edge_index, ops.GraphCSR):
        return ops.adopt_csr(edge_index)
    return edge_index


def _pad4(x, weights, chain=False):
    """Zero-pad the feature axis of ``x`` and the input axis of ``weights`` to a multiple of 4 floats (16-byte rows) so the
    vector hop kernels and the tensor-core GEMM take layers whose input width is not one (21 / 25 in everyday.json).
    The padding columns are exact zeros end to end; ``F.pad`` keeps autograd (the gradients are sliced back)."""
    pad = (-x.shape[1]) % 4
    if chain and x.shape[1] < 32 and PAD_CHAIN_WIDTH:
        pad = 32 - x.shape[1]      # one 128-byte slice: the K hops of the layer run as ONE chain launch (needs F % 32 == 0)
    if pad == 0 or not PAD_INPUT_WIDTH:
        return x, weights
    return nn.functional.pad(x, (0, pad)), [nn.functional.pad(w, (0, pad)) for w in weights]


PAD_INPUT_WIDTH = True
PAD_CHAIN_WIDTH = os.environ.get("DCB200_PAD_CHAIN", "1") == "1"   # TAGConv inputs narrower than 32 (21 / 25): pad to 32 instead of 24 / 28


# ------------------------------------------------------------------------------- TAGConv
class TAGConv(nn.Module):
    """PyG ``TAGConv(in, out, K=3, bias=True, normalize=True)``; keys ``lins.{0..K}.weight``, ``bias``.
    out = act( sum_k A_hat^k X W_k^T + b ): K hops in one chain launch (K1) + one multi-segment GEMM with bias / ReLU
    epilogue (K2); backward dH_k = dOut W_k, dW_k = dOut^T H_k, db = colsum(dOut), then the K transposed hops as one chain
    accumulating in place (``torch.ops.dcb200.tag_conv`` / ``tag_conv_backward``)."""

    def __init__(self, in_channels, out_channels, K=3, bias=True, normalize=True, precision=ops.GEMM_AUTO):
        super().__init__()
        self.in_channels, self.out_channels, self.K, self.normalize = in_channels, out_channels, K, normalize
        self.lins = nn.ModuleList([_Lin(in_channels, out_channels) for _ in range(K + 1)])
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        self.precision = precision

    def forward(self, x, edge_index, relu=False, ptr=None):
        x, ws = _pad4(x, [l.weight for l in self.lins], chain=True)
        return torch.ops.dcb200.tag_conv(x, _edge_tensor(edge_index), ws, self.bias, relu, self.normalize, self.precision, ptr)[0]


def layer_stack(layers, x, edge_index, relu=True, ptr=None, between=None):
    """``for conv in layers: x = act(conv(x, edge_index))`` (models/model.py:69-78) for TAGConv / GCNConv layers on ONE graph.
    A relabelled large graph (one cloud of >= 32768 points from ``knn_graph`` / ``radius_graph``) is entered once and left
    once: its features stay in the structure's node order from layer to layer (bit-identical to the plain loop, which
    permutes in and out of every layer).  ``between``: optional callable applied after every layer (dropout)."""
    mode = {"TAGConv": "tag", "GCNConv": "gcn"}.get(type(layers[0]).__name__) if len(layers) else None
    same = mode is not None and all(type(l) is type(layers[0]) for l in layers)
    if mode == "tag" and same and not all(getattr(l, "normalize", True) for l in layers):
        same = False
    g = None
    if same and not isinstance(edge_index, ops.GraphCSR):
        x, edge_index, g = ops.stack_enter(x, edge_index, mode, ptr)
    for conv in layers:
        x = conv(x, edge_index, relu=relu, ptr=ptr)
        if between is not None:
            x = between(x)
    return ops.stack_exit(x, g)


# ------------------------------------------------------------------------------- GCNConv
class GCNConv(nn.Module):
    """PyG ``GCNConv(in, out)`` defaults; keys ``lin.weight``, ``bias``."""

    def __init__(self, in_channels, out_channels, bias=True, precision=ops.GEMM_AUTO):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin = _Lin(in_channels, out_channels, init="glorot")
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        self.precision = precision

    def forward(self, x, edge_index, relu=False, ptr=None):
        return torch.ops.dcb200.gcn_conv(x, _edge_tensor(edge_index), self.lin.weight, self.bias, relu, self.precision, ptr)


# ------------------------------------------------------------------------------- GATConv
class GATConv(nn.Module):
    """PyG 2.5.x ``GATConv(in, out)`` defaults (heads=1, concat, slope 0.2, add_self_loops);
    keys ``lin.weight``, ``att_src``, ``att_dst``, ``bias``."""

    def __init__(self, in_channels, out_channels, heads=1, negative_slope=0.2, bias=True, precision=ops.GEMM_AUTO):
        super().__init__()
        if heads != 1:
            raise NotImplementedError("GATConv: the reference constructs heads=1 (models/model.py:45); heads>1 unsupported")
        self.in_channels, self.out_channels, self.heads, self.negative_slope = in_channels, out_channels, heads, negative_slope
        self.lin = _Lin(in_channels, heads * out_channels, init="glorot")
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        _glorot_(self.att_src)
        _glorot_(self.att_dst)
        self.bias = nn.Parameter(torch.zeros(heads * out_channels)) if bias else None
        self.precision = precision

    def forward(self, x, edge_index, relu=False, ptr=None):
        return torch.ops.dcb200.gat_conv(x, _edge_tensor(edge_index), self.lin.weight, self.att_src, self.att_dst, self.bias,
                                         float(self.negative_slope), relu, self.precision, ptr)[0]


# ------------------------------------------------------------------------------- MPNN (A9 extension)
class MPNNLayer(nn.Module):
    """``MPNNLayer(in, out)``; keys ``edge_mlp.{0,2}.{weight,bias}``, ``node_mlp.{0,2}.{weight,bias}``
    (``nn.Sequential(Linear, ReLU, Linear)`` each).  Residual when ``in == out``.  Needs ``out % 4 == 0``."""

    def __init__(self, in_channels, out_channels, precision=ops.GEMM_AUTO):
        super().__init__()
        if out_channels % 4:
            raise NotImplementedError("MPNNLayer: out_channels must be a multiple of 4")
        self.in_channels, self.out_channels, self.precision = in_channels, out_channels, precision
        self.edge_mlp = nn.Sequential(nn.Linear(2 * in_channels, out_channels), nn.ReLU(), nn.Linear(out_channels, out_channels))
        self.node_mlp = nn.Sequential(nn.Linear(in_channels + out_channels, out_channels), nn.ReLU(),
                                      nn.Linear(out_channels, out_channels))

    def forward(self, x, edge_index, relu=False, ptr=None):
        out = torch.ops.dcb200.mpnn_layer(x, _edge_tensor(edge_index), self.edge_mlp[0].weight, self.edge_mlp[0].bias,
                                          self.edge_mlp[2].weight, self.edge_mlp[2].bias, self.node_mlp[0].weight,
                                          self.node_mlp[0].bias, self.node_mlp[2].weight, self.node_mlp[2].bias,
                                          self.in_channels == self.out_channels, self.precision, ptr)[0]
        return torch.relu(out) if relu else out
