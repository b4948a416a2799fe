"""Dense layers of the decoder MLP (models/model.py:52-64,88) on the tcgen05 3xTF32 GEMM.

``linear(segs, weight, bias, relu)`` computes ``act([segs[0] | segs[1] | ...] weight^T + bias)`` without ever
materialising the concatenation: every segment is one K-segment of a single ``dc_gemm`` call (the reference
concatenates the soft-node features with the attention heads, models/model.py:82-88 — 1.5 GB at 512k nodes).
Backward: dX_s = dY W_s, dW_s = dY^T X_s (written straight into the column slice of dW), db = colsum(dY), with the
fused ReLU undone by ``dc_relu_bwd``.
"""
import torch

from . import ops

_f32 = torch.float32


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weight, bias, relu, *segs):
        segs = [s.contiguous() if s.stride(-1) != 1 else s for s in segs]
        M, N = segs[0].shape[0], weight.shape[0]
        offs, pairs, o = [], [], 0
        for s in segs:
            pairs.append((s, weight[:, o:o + s.shape[1]]))
            offs.append(o)
            o += s.shape[1]
        if o != weight.shape[1]:
            raise ValueError(f"linear: segments are {o} wide, weight expects {weight.shape[1]}")
        y = ops.gemm(pairs, M, N, trans_b=True, bias=bias, relu=relu)
        ctx.save_for_backward(weight, y if relu else None, *segs)
        ctx.relu, ctx.offs, ctx.has_bias = relu, offs, bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        weight, y, *segs = ctx.saved_tensors
        dy, db = ops.relu_bwd_db(y, dy.contiguous(), ctx.relu, ctx.has_bias)   # ReLU backward + bias gradient in one pass
        M, N = dy.shape
        dW = torch.empty_like(weight)
        dsegs = []
        for s, o, need in zip(segs, ctx.offs, ctx.needs_input_grad[3:]):
            k = s.shape[1]
            Ws = weight[:, o:o + k]
            dsegs.append(ops.gemm([(dy, Ws)], M, k, trans_b=False) if need else None)          # dY W_s
            ops.gemm([(dy, s)], N, k, trans_a=True, trans_b=False, out=dW[:, o:o + k])          # dY^T X_s
        return (dW, db, None, *dsegs)


def linear(segs, weight, bias=None, relu=False):
    """act(concat(segs, dim=1) @ weight.T + bias); ``segs`` is a list of fp32 [M, k_s] CUDA tensors."""
    return _LinearFn.apply(weight, bias, bool(relu), *segs)
