"""Restated message-passing layers (oracle; CPU PyTorch; test infrastructure only).

Reference call sites: ``models/model.py:39`` picks the class, ``:45,49`` construct
``conv_layer(in, hidden)`` with every other argument at its default, ``:71,77`` call
``conv(x, edge_index)``.  The arithmetic is torch_geometric 2.5.2's
(``nn/conv/tag_conv.py``, ``gcn_conv.py``, ``gat_conv.py``, ``message_passing.py``,
``utils/softmax.py``)  [3P, restated from the published algorithm — PARITY UNPINNED,
see oracle/__init__.py].  Parameter names and shapes are PyG's so that a reference
``state_dict`` loads (``train.py:125``, ``eval.py:36,89``).

``edge_index`` is ``int64 [2, E]``; row 0 = source j, row 1 = target i
(flow ``source_to_target``, ``utils/graph_utils.py:13``).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------- helpers
def _scatter_sum(src, index, n):
    """``torch_geometric.utils.scatter(reduce='sum')``: zeros + scatter_add_ along dim 0."""
    shape = (n,) + tuple(src.shape[1:])
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    return torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add_(0, idx, src)


def remove_self_loops(edge_index):
    mask = edge_index[0] != edge_index[1]
    return edge_index[:, mask]


def add_self_loops(edge_index, n):
    loop = torch.arange(n, dtype=edge_index.dtype, device=edge_index.device)
    return torch.cat([edge_index, loop.unsqueeze(0).repeat(2, 1)], dim=1)


def add_remaining_self_loops(edge_index, n):
    """PyG ``add_remaining_self_loops`` with ``edge_attr=None``: the non-loop edges in
    their original order followed by one self loop per node (existing loops are
    thereby kept exactly once)."""
    return add_self_loops(remove_self_loops(edge_index), n)


def gcn_norm(edge_index, num_nodes, add_self_loops_flag=False, dtype=torch.float32):
    """PyG ``gcn_norm(edge_index, None, N, improved=False, add_self_loops, 'source_to_target')``.

    TAGConv calls it with ``add_self_loops=False`` (``tag_conv.py``), GCNConv with ``True``.
    deg = in-degree at the target (duplicates and existing self loops counted);
    ``dis = deg.pow(-0.5)``, ``inf -> 0``; ``w_e = dis[row] * 1 * dis[col]`` in that order.
    """
    if add_self_loops_flag:
        edge_index = add_remaining_self_loops(edge_index, num_nodes)
    w = torch.ones(edge_index.shape[1], dtype=dtype, device=edge_index.device)
    row, col = edge_index[0], edge_index[1]
    deg = _scatter_sum(w, col, num_nodes)
    dis = deg.pow(-0.5)
    dis = dis.masked_fill(dis == float("inf"), 0.0)
    w = dis[row] * w * dis[col]
    return edge_index, w


def propagate(h, edge_index, w, num_nodes=None):
    """``MessagePassing.propagate`` with ``aggr='add'`` and
    ``message = edge_weight.view(-1, 1) * x_j``:  h'[i] = sum_{e: col_e = i} w_e * h[row_e]."""
    n = h.shape[0] if num_nodes is None else num_nodes
    x_j = h.index_select(0, edge_index[0])
    msg = w.view(-1, *([1] * (h.dim() - 1))) * x_j
    return _scatter_sum(msg, edge_index[1], n)


def _glorot_(t):
    a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        t.uniform_(-a, a)


def _kaiming_uniform_linear_(t):
    """PyG ``Linear`` default init = ``kaiming_uniform(fan=in, a=sqrt(5))`` = U(+-1/sqrt(in))."""
    bound = 1.0 / math.sqrt(t.size(-1)) if t.size(-1) > 0 else 0.0
    with torch.no_grad():
        t.uniform_(-bound, bound)


class _Lin(nn.Module):
    """Bias-free linear holding ``weight [out, in]`` (PyG ``nn.dense.linear.Linear``)."""

    def __init__(self, i, o, init="kaiming"):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(o, i))
        (_glorot_ if init == "glorot" else _kaiming_uniform_linear_)(self.weight)

    def forward(self, x):
        return F.linear(x, self.weight)


# --------------------------------------------------------------------------- TAGConv
class TAGConv(nn.Module):
    """out = sum_{k=0..K} A_hat^k X W_k^T + b, A_hat = D^-1/2 A D^-1/2 (no self loops).

    Parameters: ``lins.{0..K}.weight [out, in]``, ``bias [out]`` (zeros).  K = 3.
    """

    def __init__(self, in_channels, out_channels, K=3, bias=True, normalize=True):
        super().__init__()
        self.in_channels, self.out_channels, self.K, self.normalize = in_channels, out_channels, K, normalize
        self.lins = nn.ModuleList([_Lin(in_channels, out_channels) for _ in range(K + 1)])
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None

    def forward(self, x, edge_index):
        n = x.shape[0]
        if self.normalize:
            edge_index, w = gcn_norm(edge_index, n, False, x.dtype)
        else:
            w = torch.ones(edge_index.shape[1], dtype=x.dtype, device=x.device)
        out = self.lins[0](x)
        for lin in self.lins[1:]:
            x = propagate(x, edge_index, w, n)
            out = out + lin(x)
        if self.bias is not None:
            out = out + self.bias
        return out


# --------------------------------------------------------------------------- GCNConv
class GCNConv(nn.Module):
    """out = A_hat (X W^T) + b with A_hat built on the graph plus remaining self loops.

    Parameters: ``lin.weight [out, in]`` (glorot), ``bias [out]`` (zeros).
    """

    def __init__(self, in_channels, out_channels, bias=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin = _Lin(in_channels, out_channels, init="glorot")
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None

    def forward(self, x, edge_index):
        n = x.shape[0]
        edge_index, w = gcn_norm(edge_index, n, True, x.dtype)
        x = self.lin(x)
        out = propagate(x, edge_index, w, n)
        if self.bias is not None:
            out = out + self.bias
        return out


# --------------------------------------------------------------------------- GATConv
def segment_softmax(src, index, n):
    """``torch_geometric.utils.softmax(src, index, num_nodes=n)``."""
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    src_max = torch.full((n,) + tuple(src.shape[1:]), float("-inf"), dtype=src.dtype, device=src.device)
    src_max = src_max.scatter_reduce(0, idx, src.detach(), reduce="amax", include_self=True)
    out = (src - src_max.index_select(0, index)).exp()
    out_sum = _scatter_sum(out, index, n) + 1e-16
    return out / out_sum.index_select(0, index)


class GATConv(nn.Module):
    """PyG 2.5.x ``GATConv(in, out)`` defaults: heads=1, concat=True, negative_slope=0.2,
    dropout=0, add_self_loops=True, bias=True.

    Parameters: ``lin.weight [H*C, in]`` (glorot), ``att_src``/``att_dst [1, H, C]`` (glorot),
    ``bias [H*C]`` (zeros).
    """

    def __init__(self, in_channels, out_channels, heads=1, negative_slope=0.2, bias=True):
        super().__init__()
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.negative_slope = negative_slope
        self.lin = _Lin(in_channels, heads * out_channels, init="glorot")
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        _glorot_(self.att_src)
        _glorot_(self.att_dst)
        self.bias = nn.Parameter(torch.zeros(heads * out_channels)) if bias else None

    def forward(self, x, edge_index):
        n, H, C = x.shape[0], self.heads, self.out_channels
        xs = self.lin(x).view(n, H, C)
        a_src = (xs * self.att_src).sum(-1)
        a_dst = (xs * self.att_dst).sum(-1)
        edge_index = add_self_loops(remove_self_loops(edge_index), n)
        row, col = edge_index[0], edge_index[1]
        alpha = F.leaky_relu(a_src[row] + a_dst[col], self.negative_slope)
        alpha = segment_softmax(alpha, col, n)
        msg = alpha.unsqueeze(-1) * xs.index_select(0, row)
        out = _scatter_sum(msg, col, n).reshape(n, H * C)
        if self.bias is not None:
            out = out + self.bias
        return out


# --------------------------------------------------------------------------- MPNN (extension)
class MPNNLayer(nn.Module):
    """north_star's "edge-MLP update / scatter-sum / node-MLP update with residual" layer.

    **No reference counterpart**  — this definition IS the spec:
        m_e  = W_e2 relu(W_e1 [x_i || x_j] + b_e1) + b_e2        (i = target, j = source)
        a_i  = sum_{e: col_e = i} m_e
        x'_i = (x_i if in == out else 0) + W_n2 relu(W_n1 [x_i || a_i] + b_n1) + b_n2
    Parity status: unpinned (self-defined oracle).
    """

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.edge_mlp = nn.Sequential(nn.Linear(2 * in_channels, out_channels), nn.ReLU(),
                                      nn.Linear(out_channels, out_channels))
        self.node_mlp = nn.Sequential(nn.Linear(in_channels + out_channels, out_channels), nn.ReLU(),
                                      nn.Linear(out_channels, out_channels))

    def forward(self, x, edge_index):
        row, col = edge_index[0], edge_index[1]
        m = self.edge_mlp(torch.cat([x[col], x[row]], dim=-1))
        a = _scatter_sum(m, col, x.shape[0])
        upd = self.node_mlp(torch.cat([x, a], dim=-1))
        return x + upd if self.in_channels == self.out_channels else upd


# --------------------------------------------------------------------------- independent checks
def tag_dense_fp64(x, edge_index, weights, bias):
    """Second, independent formulation used to pin the restatement: dense fp64
    ``sum_k A_hat^k X W_k^T + b`` with A_hat[i, j] = (#edges j->i) / sqrt(deg_i deg_j)."""
    n = x.shape[0]
    A = torch.zeros(n, n, dtype=torch.float64)
    A.index_put_((edge_index[1], edge_index[0]), torch.ones(edge_index.shape[1], dtype=torch.float64),
                 accumulate=True)
    deg = A.sum(1)
    dis = torch.where(deg > 0, deg.pow(-0.5), torch.zeros_like(deg))
    A = dis[:, None] * A * dis[None, :]
    h = x.double()
    out = h @ weights[0].double().t()
    for W in weights[1:]:
        h = A @ h
        out = out + h @ W.double().t()
    return out + bias.double()
