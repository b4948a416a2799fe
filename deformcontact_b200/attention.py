"""N1: the dense cross attention of models/model.py:7-21 on the tcgen05 3xTF32 GEMMs.

Per head (one shared ``nn.Linear`` L for queries and keys, no 1/sqrt(d) scale, values = raw collider features):

    scores = L(x_resting) L(x_rigid)^T,   attn = softmax(scores, dim=-1),   out = attn x_rigid

Everything runs in libdcb200 through the C ABI: the two projections are one GEMM each over all nodes of the
batch; inside every attention group (``attn_group`` graphs, see model.py) the score matrix, its softmax and the
two products with it are ``dc_gemm_batched`` (one launch of the tcgen05 kind::tf32 kernel with the 3-term split per
product type over ALL groups, fp32-class accuracy) and ``dc_softmax_rows`` / ``dc_softmax_bwd_rows``.  Backward is hand-derived (one ``autograd.Function`` over all
heads): dP = dO Xr^T, dXr = P^T dO, dS = P o (dP - rowsum(dP o P)), dQ = dS K, dK = dS^T Q, then the projection
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


class _Blocks(list):
    """The per-group [ns_g, nr_g] matrices of one head; ``whole`` = the [sum ns_g, nr] matrix they are row blocks of when every
    group has the same number of collider nodes (the regular case: equal collider meshes, ``attn_group`` graphs per group)."""
    whole = None


def _group_matrices(live, dev):
    """Uninitialised score / weight matrices of the live groups, rows 16-byte aligned (TMA operands of the batched products)."""
    widths = {r1 - r0 for _, _, r0, r1 in live}
    out = _Blocks()
    if len(widths) == 1:
        nr = widths.pop()
        ld = (nr + 3) // 4 * 4
        whole = torch.empty((sum(s1 - s0 for s0, s1, _, _ in live), ld), dtype=_f32, device=dev)[:, :nr]
        row = 0
        for s0, s1, _, _ in live:
            out.append(whole[row:row + s1 - s0])
            row += s1 - s0
        out.whole = whole
    else:
        out.extend(torch.empty((s1 - s0, (r1 - r0 + 3) // 4 * 4), dtype=_f32, device=dev)[:, :r1 - r0] for s0, s1, r0, r1 in live)
    return out


class _AttnFn(torch.autograd.Function):
    """All heads at once: every product type is ONE batched launch over heads x groups."""

    @staticmethod
    def forward(ctx, xs, xr, groups, *params):   # params = W_0, b_0, W_1, b_1, ...
        xs, xr = xs.contiguous(), xr.contiguous()
        Ns, F = xs.shape
        Nr = xr.shape[0]
        H = len(params) // 2
        dev = xs.device
        live = []
        outs = [torch.empty((Ns, F), dtype=_f32, device=dev) for _ in range(H)]
        for s0, s1, r0, r1 in groups:
            if s1 - s0 == 0:
                continue
            if r1 - r0 == 0:   # softmax over an empty set: the reference's empty mm gives zeros
                for o in outs:
                    o[s0:s1].zero_()
                continue
            live.append((s0, s1, r0, r1))
        qs, ks, probs = [], [], []
        for h in range(H):
            W, b = params[2 * h], params[2 * h + 1]
            qs.append(ops.gemm([(xs, W)], Ns, F, trans_b=True, bias=b))     # L_h(x_resting)  [Ns, F]
            ks.append(ops.gemm([(xr, W)], Nr, F, trans_b=True, bias=b))     # L_h(x_rigid)    [Nr, F]
            probs.append(_group_matrices(live, dev))
        HG = [(h, i, g) for h in range(H) for i, g in enumerate(live)]
        ops.gemm_batched([(qs[h][s0:s1], ks[h][r0:r1], probs[h][i]) for h, i, (s0, s1, r0, r1) in HG], trans_b=True)      # scores
        for h in range(H):
            whole = getattr(probs[h], "whole", None)
            if whole is not None:     # groups of equal width are row blocks of one matrix: one softmax launch per head
                softmax_rows_(whole)
            else:
                for P in probs[h]:
                    softmax_rows_(P)
        ops.gemm_batched([(probs[h][i], xr[r0:r1], outs[h][s0:s1]) for h, i, (s0, s1, r0, r1) in HG], trans_b=False)      # attn @ x_rigid
        ctx.save_for_backward(xs, xr, *params, *qs, *ks, *outs)
        ctx.probs, ctx.groups, ctx.H = probs, live, H
        return tuple(outs)

    @staticmethod
    def backward(ctx, *douts):
        H, live, probs = ctx.H, ctx.groups, ctx.probs
        saved = ctx.saved_tensors
        xs, xr = saved[0], saved[1]
        params = saved[2:2 + 2 * H]
        qs, ks, outs = saved[2 + 2 * H:2 + 3 * H], saved[2 + 3 * H:2 + 4 * H], saved[2 + 4 * H:2 + 5 * H]
        Ns, F = xs.shape
        Nr = xr.shape[0]
        dev = xs.device
        douts = [d.contiguous() for d in douts]
        # every row of dq / dk / dxr_h inside a live group is written by the batched products below; only rows of groups that were
        # skipped in forward (no resting or no rigid nodes) need zeros — none in a regular batch, so no full-size fills
        cov_s = sum(s1 - s0 for s0, s1, _, _ in live) == Ns
        cov_r = sum(r1 - r0 for _, _, r0, r1 in live) == Nr
        new_s = torch.empty if cov_s else torch.zeros
        new_r = torch.empty if cov_r else torch.zeros
        dq = [new_s((Ns, F), dtype=_f32, device=dev) for _ in range(H)]
        dk = [new_r((Nr, F), dtype=_f32, device=dev) for _ in range(H)]
        dxr_h = [new_r((Nr, F), dtype=_f32, device=dev) for _ in range(H)]
        dPs = [_group_matrices(live, dev) for _ in range(H)]
        HG = [(h, i, g) for h in range(H) for i, g in enumerate(live)]
        # softmax backward dS = P o (dP - rowsum(dP o P)) fused into the epilogue of dP = dO Xr^T: the row sums equal
        # rowdot(dO, O) because O = P Xr, so they are known before the product starts
        D = [ops.rowdot(douts[h], outs[h]) for h in range(H)]
        ops.gemm_batched([(douts[h][s0:s1], xr[r0:r1], dPs[h][i], probs[h][i], D[h][s0:s1]) for h, i, (s0, s1, r0, r1) in HG],
                         trans_b=True)                                                                                   # dS
        ops.gemm_batched([(probs[h][i], douts[h][s0:s1], dxr_h[h][r0:r1]) for h, i, (s0, s1, r0, r1) in HG],
                         trans_a=True, trans_b=False)                                                                    # dXr_h = P^T dO
        ops.gemm_batched([(dPs[h][i], ks[h][r0:r1], dq[h][s0:s1]) for h, i, (s0, s1, r0, r1) in HG], trans_b=False)      # dQ = dS K
        ops.gemm_batched([(dPs[h][i], qs[h][s0:s1], dk[h][r0:r1]) for h, i, (s0, s1, r0, r1) in HG],
                         trans_a=True, trans_b=False)                                                                    # dK = dS^T Q
        ctx.probs = None
        # projections: q = xs W^T + b, k = xr W^T + b
        dxs, dxr, grads = None, None, []
        for h in range(H):
            W = params[2 * h]
            if dxs is None:
                dxs = ops.gemm([(dq[h], W)], Ns, F, trans_b=False)
                dxr = dxr_h[h]
            else:
                ops.gemm([(dq[h], W)], Ns, F, trans_b=False, out=dxs, accumulate=True)
                dxr = dxr + dxr_h[h]
            ops.gemm([(dk[h], W)], Nr, F, trans_b=False, out=dxr, accumulate=True)
            dW = ops.gemm([(dq[h], xs)], F, F, trans_a=True, trans_b=False)
            ops.gemm([(dk[h], xr)], F, F, trans_a=True, trans_b=False, out=dW, accumulate=True)
            grads += [dW, ops.colsum(dq[h]) + ops.colsum(dk[h])]
        return (dxs, dxr, None, *grads)


def cross_attention(x_resting, x_rigid, heads, ptr_s, ptr_r, group=None, concat=True):
    """``heads``: iterable of ``nn.Linear``; returns the head outputs concatenated along the feature axis
    (models/model.py:14-21), or as a list when ``concat`` is False (the decoder consumes them as GEMM K-segments).
    ``ptr_s`` / ``ptr_r``: host lists of graph offsets of the two batches."""
    groups = _groups(ptr_s, ptr_r, group)
    params = []
    for h in heads:
        params += [h.weight, h.bias]
    outs = list(_AttnFn.apply(x_resting, x_rigid, groups, *params))
    return torch.cat(outs, dim=-1) if concat else outs
