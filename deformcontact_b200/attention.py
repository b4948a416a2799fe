"""N1: the dense cross attention of models/model.py:7-21 on the tcgen05 3xTF32 GEMMs.

Per head (one shared ``nn.Linear`` L for queries and keys, no 1/sqrt(d) scale, values = raw collider features):

    scores = L(x_resting) L(x_rigid)^T,   attn = softmax(scores, dim=-1),   out = attn x_rigid

Everything runs in libdcb200 through the C ABI: the two projections are one GEMM each over all nodes of the
batch; inside every attention group (``attn_group`` graphs, see model.py) the score matrix, its softmax and the
two products with it are ``dc_gemm_batched`` (one launch of the tcgen05 kind::tf32 kernel with the 3-term split per
product type over ALL groups, fp32-class accuracy) and ``dc_softmax_rows`` / ``dc_softmax_bwd_rows``.  Backward is hand-derived (one ``autograd.Function`` per
head): dP = dO Xr^T, dXr = P^T dO, dS = P o (dP - rowsum(dP o P)), dQ = dS K, dK = dS^T Q, then the projection
gradients over all nodes at once.  The softmax backward is fused into the epilogue of the dP product
(``rowsum(dP o P) = rowdot(dO, O)``).  The attention weights of a group are kept for backward ([ns, nr] fp32).
"""
import torch

from . import _abi, ops

_f32 = torch.float32


def softmax_rows_(S):
    """In-place row softmax of a contiguous-row fp32 matrix; see dc_softmax_rows."""
    _abi.call("dc_softmax_rows", ops._ptr(S), ops._rows(S, "S"), S.shape[0], S.shape[1], ops._stream())
    return S


def softmax_bwd_rows_(P, dP):
    """In place on dP: dS = P o (dP - rowsum(dP o P)); see dc_softmax_bwd_rows."""
    _abi.call("dc_softmax_bwd_rows", ops._ptr(P), ops._rows(P, "P"), ops._ptr(dP), ops._rows(dP, "dP"), P.shape[0], P.shape[1],
              ops._stream())
    return dP


def _groups(ptr_s, ptr_r, group):
    """[(s0, s1, r0, r1)] row ranges of the attention groups (``group`` graphs each; None = the whole batch)."""
    B = len(ptr_s) - 1
    G = B if group is None else int(group)
    return [(ptr_s[a], ptr_s[min(a + G, B)], ptr_r[a], ptr_r[min(a + G, B)]) for a in range(0, B, G)]


class _AttnHeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xs, xr, W, b, groups):
        xs, xr = xs.contiguous(), xr.contiguous()
        Ns, F = xs.shape
        Nr = xr.shape[0]
        q = ops.gemm([(xs, W)], Ns, F, trans_b=True, bias=b)     # L(x_resting)  [Ns, F]
        k = ops.gemm([(xr, W)], Nr, F, trans_b=True, bias=b)     # L(x_rigid)    [Nr, F]
        out = torch.empty((Ns, F), dtype=_f32, device=xs.device)
        live, probs = [], []
        for s0, s1, r0, r1 in groups:
            ns, nr = s1 - s0, r1 - r0
            if ns == 0:
                continue
            if nr == 0:   # softmax over an empty set: the reference's empty mm gives zeros
                out[s0:s1].zero_()
                continue
            live.append((s0, s1, r0, r1))
            probs.append(torch.empty((ns, (nr + 3) // 4 * 4), dtype=_f32, device=xs.device)[:, :nr])   # 16-byte aligned rows
        # one batched tensor-core launch per product type over all groups
        ops.gemm_batched([(q[s0:s1], k[r0:r1], P) for (s0, s1, r0, r1), P in zip(live, probs)], trans_b=True)       # scores
        for P in probs:
            softmax_rows_(P)
        ops.gemm_batched([(P, xr[r0:r1], out[s0:s1]) for (s0, s1, r0, r1), P in zip(live, probs)], trans_b=False)   # attn @ x_rigid
        ctx.save_for_backward(xs, xr, W, q, k, out)
        ctx.probs, ctx.groups = probs, live
        return out

    @staticmethod
    def backward(ctx, dout):
        xs, xr, W, q, k, out = ctx.saved_tensors
        dout = dout.contiguous()
        Ns, F = xs.shape
        Nr = xr.shape[0]
        groups, probs = ctx.groups, ctx.probs
        dq = torch.zeros((Ns, F), dtype=_f32, device=xs.device)
        dk = torch.zeros((Nr, F), dtype=_f32, device=xs.device)
        dxr = torch.zeros((Nr, F), dtype=_f32, device=xs.device)
        dPs = [torch.empty((P.shape[0], P.stride(0)), dtype=_f32, device=xs.device)[:, :P.shape[1]] for P in probs]
        G = list(zip(groups, probs, dPs))
        # softmax backward dS = P o (dP - rowsum(dP o P)) fused into the epilogue of dP = dO Xr^T: the row sums equal
        # rowdot(dO, O) because O = P Xr, so they are known before the product starts
        D = ops.rowdot(dout, out)
        ops.gemm_batched([(dout[s0:s1], xr[r0:r1], dP, P, D[s0:s1]) for (s0, s1, r0, r1), P, dP in G], trans_b=True)          # dS
        ops.gemm_batched([(P, dout[s0:s1], dxr[r0:r1]) for (s0, s1, r0, r1), P, dP in G], trans_a=True, trans_b=False)       # dXr = P^T dO
        ops.gemm_batched([(dP, k[r0:r1], dq[s0:s1]) for (s0, s1, r0, r1), P, dP in G], trans_b=False)                         # dQ = dS K
        ops.gemm_batched([(dP, q[s0:s1], dk[r0:r1]) for (s0, s1, r0, r1), P, dP in G], trans_a=True, trans_b=False)           # dK = dS^T Q
        ctx.probs = None
        # projections: q = xs W^T + b, k = xr W^T + b
        dxs = ops.gemm([(dq, W)], Ns, F, trans_b=False)
        ops.gemm([(dk, W)], Nr, F, trans_b=False, out=dxr, accumulate=True)
        dW = ops.gemm([(dq, xs)], F, F, trans_a=True, trans_b=False)
        ops.gemm([(dk, xr)], F, F, trans_a=True, trans_b=False, out=dW, accumulate=True)
        db = ops.colsum(dq) + ops.colsum(dk)
        return dxs, dxr, dW, db, None


def cross_attention(x_resting, x_rigid, heads, ptr_s, ptr_r, group=None, concat=True):
    """``heads``: iterable of ``nn.Linear``; returns the head outputs concatenated along the feature axis
    (models/model.py:14-21), or as a list when ``concat`` is False (the decoder consumes them as GEMM K-segments).
    ``ptr_s`` / ``ptr_r``: host lists of graph offsets of the two batches."""
    groups = _groups(ptr_s, ptr_r, group)
    outs = [_AttnHeadFn.apply(x_resting, x_rigid, h.weight, h.bias, groups) for h in heads]
    return torch.cat(outs, dim=-1) if concat else outs
