"""Drop-in message-passing layers: ``TAGConv`` / ``GCNConv`` / ``GATConv`` (+ ``MPNNLayer``).

Same constructor signature ``(in_channels, out_channels)``, call ``layer(x, edge_index)`` and
state-dict keys as torch_geometric 2.5.2, so they slot into the reference's
``GraphNet.__init__`` / ``forward`` (models/model.py:39-50, 69-78) and load its checkpoints
(train.py:125, eval.py:36,89).  Forward and backward run entirely in libdcb200 (hand-written
sm_100a CUDA reached through the C ABI); autograd sees one ``torch.autograd.Function`` per layer.

``edge_index`` may be the usual ``int64 [2, E]`` tensor (the CSR pair is built once per tensor
and cached) or a prebuilt ``ops.GraphCSR``.
"""
import math

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


# True: modules dispatch through the registered ``torch.ops.dcb200.*`` custom ops (torch_ops.py);
# False: through the equivalent ``torch.autograd.Function``s below.  Same kernels either way.
USE_TORCH_OPS = True


def _structure(edge_index, n, mode, ptr=None):
    return ops.graph_csr(edge_index, n, mode, ptr)


def _pad4(x, weights):
    """Zero-pad the feature axis of ``x`` and the input axis of ``weights`` to a multiple of 4 floats (16-byte rows) so the
    vector hop kernels and the tensor-core GEMM take layers whose input width is not one (21 / 25 in everyday.json).
    The padding columns are exact zeros end to end; ``F.pad`` keeps autograd (the gradients are sliced back)."""
    pad = (-x.shape[1]) % 4
    if pad == 0 or not PAD_INPUT_WIDTH:
        return x, weights
    return nn.functional.pad(x, (0, pad)), [nn.functional.pad(w, (0, pad)) for w in weights]


PAD_INPUT_WIDTH = True


# ------------------------------------------------------------------------------- TAGConv
class _TAGConvFn(torch.autograd.Function):
    """out = act( sum_k A_hat^k X W_k^T + b ).  Forward: K hops (dc_spmm) + one multi-segment GEMM
    with bias/ReLU epilogue.  Backward: dH_k = dOut W_k, dW_k = dOut^T H_k (split-K, fixed order),
    db = colsum(dOut), then K transposed hops with fused accumulate:
    g_K = dH_K, g_{k-1} = dH_{k-1} + A_hat^T g_k, dX = g_0."""

    @staticmethod
    def forward(ctx, x, g, bias, relu, precision, *weights):
        x = x.contiguous()
        N, Fi = x.shape
        K = len(weights) - 1
        Fo = weights[0].shape[0]
        hs = [g.to_internal(x)]        # a relabelled large graph (ops.REORDER) runs the whole layer in its own node order
        if K > 0:
            buf = torch.empty((N, K * Fi), dtype=x.dtype, device=x.device)
            for k in range(K):
                hs.append(buf[:, k * Fi:(k + 1) * Fi])
            ops.propagate_chain(g, [(hs[k], None, hs[k + 1]) for k in range(K)], internal=True)   # h_{k+1} = A_hat h_k
        out = ops.gemm([(h, w) for h, w in zip(hs, weights)], N, Fo, False, True, bias=bias, relu=relu,
                       precision=precision)
        out = g.from_internal(out)
        ctx.g, ctx.relu, ctx.precision, ctx.has_bias = g, relu, precision, bias is not None
        ctx.save_for_backward(out if relu else None, *hs, *weights)
        return out

    @staticmethod
    def backward(ctx, dout):
        saved = ctx.saved_tensors
        out = saved[0]
        n = (len(saved) - 1) // 2
        hs, weights = saved[1:1 + n], saved[1 + n:]
        g, K = ctx.g, n - 1
        dout = dout.contiguous()
        if ctx.relu:
            dout = ops.relu_bwd(out, dout)
        dout = g.to_internal(dout)     # hs are saved in the structure's node order
        N, Fo = dout.shape
        Fi = hs[0].shape[1]
        need_x = ctx.needs_input_grad[0]
        dws = [None] * n
        for k in range(n):
            if ctx.needs_input_grad[5 + k]:
                dws[k] = ops.gemm([(dout, hs[k])], Fo, Fi, True, False, precision=ctx.precision)
        db = ops.colsum(dout) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        dx = None
        if need_x:
            gk = ops.gemm([(dout, weights[K])], N, Fi, False, False, precision=ctx.precision)
            if ops.K1_CHAIN >= 2 and K > 0:
                # all dH_k first, then the transposed hops as one chain, accumulating in place: dH_{k-1} += A_hat^T g_k
                dhs = [ops.gemm([(dout, weights[k])], N, Fi, False, False, precision=ctx.precision) for k in range(K)] + [gk]
                ops.propagate_chain(g, [(dhs[k + 1], dhs[k], dhs[k]) for k in range(K - 1, -1, -1)], transpose=True, internal=True)
                gk = dhs[0]
            else:
                for k in range(K - 1, -1, -1):
                    dhk = ops.gemm([(dout, weights[k])], N, Fi, False, False, precision=ctx.precision)
                    gk = g.propagate(gk, transpose=True, add=dhk, internal=True)
            dx = g.from_internal(gk)
        return (dx, None, db, None, None, *dws)


class TAGConv(nn.Module):
    """PyG ``TAGConv(in, out, K=3, bias=True, normalize=True)``; keys ``lins.{0..K}.weight``, ``bias``."""

    def __init__(self, in_channels, out_channels, K=3, bias=True, normalize=True, precision=ops.GEMM_AUTO):
        super().__init__()
        self.in_channels, self.out_channels, self.K, self.normalize = in_channels, out_channels, K, normalize
        self.lins = nn.ModuleList([_Lin(in_channels, out_channels) for _ in range(K + 1)])
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        self.precision = precision

    def forward(self, x, edge_index, relu=False, ptr=None):
        x, ws = _pad4(x, [l.weight for l in self.lins])
        if USE_TORCH_OPS and isinstance(edge_index, torch.Tensor):
            return torch.ops.dcb200.tag_conv(x, edge_index, ws, self.bias, relu, self.normalize, self.precision, ptr)[0]
        g = _structure(edge_index, x.shape[0], "tag" if self.normalize else "plain", ptr)
        return _TAGConvFn.apply(x, g, self.bias, relu, self.precision, *ws)


# ------------------------------------------------------------------------------- GCNConv
class _GCNConvFn(torch.autograd.Function):
    """out = act( A_hat (X W^T) + b ), A_hat with remaining self loops (appended after the edges)."""

    @staticmethod
    def forward(ctx, x, g, weight, bias, relu, precision):
        x = x.contiguous()
        N = x.shape[0]
        Fo = weight.shape[0]
        xw = ops.gemm([(x, weight)], N, Fo, False, True, precision=precision)
        out = g.propagate(xw, bias=bias, relu=relu)
        ctx.g, ctx.relu, ctx.precision, ctx.has_bias = g, relu, precision, bias is not None
        ctx.save_for_backward(out if relu else None, x, weight)
        return out

    @staticmethod
    def backward(ctx, dout):
        out, x, weight = ctx.saved_tensors
        g = ctx.g
        dout = dout.contiguous()
        if ctx.relu:
            dout = ops.relu_bwd(out, dout)
        N, Fo = dout.shape
        Fi = x.shape[1]
        db = ops.colsum(dout) if (ctx.has_bias and ctx.needs_input_grad[3]) else None
        dxw = g.propagate(dout, transpose=True)
        dx = ops.gemm([(dxw, weight)], N, Fi, False, False, precision=ctx.precision) if ctx.needs_input_grad[0] else None
        dw = ops.gemm([(dxw, x)], Fo, Fi, True, False, precision=ctx.precision) if ctx.needs_input_grad[2] else None
        return dx, None, dw, db, None, None


class GCNConv(nn.Module):
    """PyG ``GCNConv(in, out)`` defaults; keys ``lin.weight``, ``bias``."""

    def __init__(self, in_channels, out_channels, bias=True, precision=ops.GEMM_AUTO):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin = _Lin(in_channels, out_channels, init="glorot")
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        self.precision = precision

    def forward(self, x, edge_index, relu=False, ptr=None):
        if USE_TORCH_OPS and isinstance(edge_index, torch.Tensor):
            return torch.ops.dcb200.gcn_conv(x, edge_index, self.lin.weight, self.bias, relu, self.precision, ptr)
        g = _structure(edge_index, x.shape[0], "gcn", ptr)
        return _GCNConvFn.apply(x, g, self.lin.weight, self.bias, relu, self.precision)


# ------------------------------------------------------------------------------- GATConv
class _GATConvFn(torch.autograd.Function):
    """heads = 1.  xs = X W^T; a_s = xs.att_src, a_d = xs.att_dst; per receiver softmax over
    leaky_relu(a_s[j] + a_d[i]) incl. the appended self loop; out = sum alpha xs[j] + b."""

    @staticmethod
    def forward(ctx, x, g, weight, att_src, att_dst, bias, slope, relu, precision):
        x = x.contiguous()
        N = x.shape[0]
        C_ = weight.shape[0]
        xs = ops.gemm([(x, weight)], N, C_, False, True, precision=precision)
        a_src, a_dst = ops.gat_scores(xs, att_src.reshape(-1), att_dst.reshape(-1), 1, C_)
        alpha_e, alpha_s = ops.gat_softmax(g, a_src, a_dst, slope)
        out = ops.spmm(g.rowptr, g.nbr, xs, edge_w=alpha_e, edge_w_index=g.eid, self_w=alpha_s, self_loop=True,
                       bias=bias, relu=relu)
        ctx.g, ctx.slope, ctx.relu, ctx.precision, ctx.has_bias = g, slope, relu, precision, bias is not None
        ctx.save_for_backward(out if relu else None, x, weight, att_src, att_dst, xs, a_src, a_dst, alpha_e, alpha_s)
        return out

    @staticmethod
    def backward(ctx, dout):
        out, x, weight, att_src, att_dst, xs, a_src, a_dst, alpha_e, alpha_s = ctx.saved_tensors
        g = ctx.g
        dout = dout.contiguous()
        if ctx.relu:
            dout = ops.relu_bwd(out, dout)
        N, C_ = dout.shape
        Fi = x.shape[1]
        db = ops.colsum(dout) if ctx.has_bias else None
        rpt, nbt, eidt = g.t
        # through the aggregation: dxs[j] = sum_{e: src=j} alpha_e dout[dst_e] + alpha_self[j] dout[j]
        dz_e, dz_s, da_dst = ops.gat_bwd_edge(g, a_src, a_dst, ctx.slope, alpha_e, alpha_s, xs, dout)
        da_src = ops.segment_sum(rpt, eidt, dz_e, dz_s, N)
        # through the scores: dxs += da_src (x) att_src + da_dst (x) att_dst  == [da_src da_dst] @ [att_src; att_dst]
        da = torch.stack([da_src, da_dst], 1).contiguous()                       # [N, 2]
        att = torch.cat([att_src.reshape(1, -1), att_dst.reshape(1, -1)], 0).contiguous()  # [2, C]
        dxs0 = ops.gemm([(da, att)], N, C_, False, False, precision=ops.GEMM_FP32)
        dxs = ops.spmm(rpt, nbt, dout, edge_w=alpha_e, edge_w_index=eidt, self_w=alpha_s, self_loop=True, add=dxs0)
        datt = ops.gemm([(da, xs)], 2, C_, True, False, precision=ops.GEMM_FP32)  # [2, C] = da^T xs
        dx = ops.gemm([(dxs, weight)], N, Fi, False, False, precision=ctx.precision) if ctx.needs_input_grad[0] else None
        dw = ops.gemm([(dxs, x)], C_, Fi, True, False, precision=ctx.precision)
        return (dx, None, dw, datt[0].reshape(att_src.shape), datt[1].reshape(att_dst.shape), db, None, None, None)


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
        g = _structure(edge_index, x.shape[0], "gat", ptr)
        return _GATConvFn.apply(x, g, self.lin.weight, self.att_src, self.att_dst, self.bias, self.negative_slope, relu,
                                self.precision)


# ------------------------------------------------------------------------------- MPNN (A9 extension)
class _MPNNFn(torch.autograd.Function):
    """Edge-MLP / scatter-sum / node-MLP residual layer (north_star wording; no reference symbol):
        m_e = W_e2 relu(W_e1 [x_i || x_j] + b_e1) + b_e2,  a_i = sum_{e: dst = i} m_e,
        x'_i = x_i + W_n2 relu(W_n1 [x_i || a_i] + b_n1) + b_n2.
    The first edge Linear is split into two NODE-level GEMMs (u = x W_e1[:, :F]^T + b, v = x W_e1[:, F:]^T) and,
    because the aggregation is a sum, the second edge Linear moves outside it:
        a = (sum_e relu(u_i + v_j)) W_e2^T + deg * b_e2.
    Per-edge work is one fused gather (dc_edge_relu); edge features never exist in HBM."""

    @staticmethod
    def forward(ctx, x, g, We1, be1, We2, be2, Wn1, bn1, Wn2, bn2, residual, precision):
        x = x.contiguous()
        N, Fi = x.shape
        Fo = We2.shape[0]
        u = ops.gemm([(x, We1[:, :Fi])], N, Fo, False, True, bias=be1, precision=precision)
        v = ops.gemm([(x, We1[:, Fi:])], N, Fo, False, True, precision=precision)
        s = ops.edge_relu(g.rowptr, g.nbr, u, v, mode=0)
        deg = (g.rowptr[1:] - g.rowptr[:-1]).to(x.dtype).unsqueeze(1)
        a = ops.gemm([(s, We2)], N, Fo, False, True, precision=precision)
        a.addcmul_(deg, be2.unsqueeze(0))
        h1 = ops.gemm([(x, Wn1[:, :Fi]), (a, Wn1[:, Fi:])], N, Fo, False, True, bias=bn1, relu=True, precision=precision)
        if residual:
            out = x.clone()
            ops.gemm([(h1, Wn2)], N, Fo, False, True, bias=bn2, out=out, accumulate=True, precision=precision)
        else:
            out = ops.gemm([(h1, Wn2)], N, Fo, False, True, bias=bn2, precision=precision)
        ctx.g, ctx.residual, ctx.precision = g, residual, precision
        ctx.save_for_backward(x, u, v, s, a, h1, deg, We1, We2, Wn1, Wn2)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, u, v, s, a, h1, deg, We1, We2, Wn1, Wn2 = ctx.saved_tensors
        g, P = ctx.g, ctx.precision
        dout = dout.contiguous()
        N, Fi = x.shape
        Fo = We2.shape[0]
        G = ops.gemm
        dh1 = ops.relu_bwd(h1, G([(dout, Wn2)], N, Fo, False, False, precision=P))
        dWn2, dbn2 = G([(dout, h1)], Fo, Fo, True, False, precision=P), ops.colsum(dout)
        dWn1 = torch.empty_like(Wn1)
        G([(dh1, x)], Fo, Fi, True, False, out=dWn1[:, :Fi], precision=P)
        G([(dh1, a)], Fo, Fo, True, False, out=dWn1[:, Fi:], precision=P)
        dbn1 = ops.colsum(dh1)
        dx = dout.clone() if ctx.residual else torch.zeros_like(x)
        G([(dh1, Wn1[:, :Fi])], N, Fi, False, False, out=dx, accumulate=True, precision=P)
        da = G([(dh1, Wn1[:, Fi:])], N, Fo, False, False, precision=P)
        dWe2 = G([(da, s)], Fo, Fo, True, False, precision=P)
        dbe2 = ops.colsum(da * deg)
        ds = G([(da, We2)], N, Fo, False, False, precision=P)
        du = ops.edge_relu(g.rowptr, g.nbr, u, v, ds, mode=1)
        rpt, nbt, _ = g.t
        dv = ops.edge_relu(rpt, nbt, v, u, ds, mode=2)
        dWe1 = torch.empty_like(We1)
        G([(du, x)], Fo, Fi, True, False, out=dWe1[:, :Fi], precision=P)
        G([(dv, x)], Fo, Fi, True, False, out=dWe1[:, Fi:], precision=P)
        dbe1 = ops.colsum(du)
        G([(du, We1[:, :Fi])], N, Fi, False, False, out=dx, accumulate=True, precision=P)
        G([(dv, We1[:, Fi:])], N, Fi, False, False, out=dx, accumulate=True, precision=P)
        return (dx if ctx.needs_input_grad[0] else None, None, dWe1, dbe1, dWe2, dbe2, dWn1, dbn1, dWn2, dbn2, None, None)


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
        g = _structure(edge_index, x.shape[0], "plain", ptr)
        out = _MPNNFn.apply(x, g, self.edge_mlp[0].weight, self.edge_mlp[0].bias, self.edge_mlp[2].weight, self.edge_mlp[2].bias,
                            self.node_mlp[0].weight, self.node_mlp[0].bias, self.node_mlp[2].weight, self.node_mlp[2].bias,
                            self.in_channels == self.out_channels, self.precision)
        return torch.relu(out) if relu else out
