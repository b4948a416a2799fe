"""``torch.ops.dcb200.*`` — the kernels as registered PyTorch custom ops.

Thin layer over ``ops.py`` (which is itself a thin layer over the C ABI): every op has a CUDA
implementation that enqueues libdcb200 kernels on the current stream, a fake (meta) implementation so
it traces under ``torch.compile`` / CUDA-graph capture tooling, and — for the layer ops — a registered
autograd formula whose backward is again a ``dcb200`` op.  The ``nn.Module``s in ``layers.py`` call
these ops.
"""
from typing import List, Optional, Tuple

import torch
from torch import Tensor

from . import ops

_lib = torch.library


def _csr(edge_index, n, mode, ptr):
    return ops.graph_csr(edge_index, n, mode, list(ptr) if ptr is not None else None)


# ----------------------------------------------------------------------------------- primitives
@_lib.custom_op("dcb200::csr_build", mutates_args=())
def csr_build(edge_index: Tensor, num_nodes: int, group_by: int, drop_self_loops: bool) -> Tuple[Tensor, Tensor, Tensor]:
    return ops.csr_build(edge_index, num_nodes, group_by, drop_self_loops)


@csr_build.register_fake
def _(edge_index, num_nodes, group_by, drop_self_loops):
    E = edge_index.shape[1]
    i32 = dict(dtype=torch.int32, device=edge_index.device)
    return torch.empty(num_nodes + 1, **i32), torch.empty(E, **i32), torch.empty(E, **i32)


@_lib.custom_op("dcb200::propagate", mutates_args=())
def propagate(h: Tensor, edge_index: Tensor, mode: str, transpose: bool, add: Optional[Tensor], bias: Optional[Tensor],
              relu: bool, ptr: Optional[List[int]]) -> Tensor:
    """One hop: act(add + A_hat h + bias) (A_hat^T if transpose); A_hat per mode 'tag' | 'gcn' | 'plain'."""
    return _csr(edge_index, h.shape[0], mode, ptr).propagate(h.contiguous(), transpose=transpose, add=add, bias=bias, relu=relu)


@propagate.register_fake
def _(h, edge_index, mode, transpose, add, bias, relu, ptr):
    return torch.empty_like(h, memory_format=torch.contiguous_format)


@_lib.custom_op("dcb200::linear", mutates_args=())
def linear(x: Tensor, weight: Tensor, bias: Optional[Tensor], relu: bool) -> Tensor:
    """act(x W^T + b) on the tcgen05 3xTF32 / fp32 GEMM (dc_gemm)."""
    return ops.gemm([(x.contiguous(), weight)], x.shape[0], weight.shape[0], False, True, bias=bias, relu=relu)


@linear.register_fake
def _(x, weight, bias, relu):
    return x.new_empty((x.shape[0], weight.shape[0]))


@_lib.custom_op("dcb200::knn_table", mutates_args=())
def knn_table(pos: Tensor, k: int, ptr: Optional[Tensor], loop: bool) -> Tensor:
    return ops.knn_table(pos, k, ptr=ptr, loop=loop)


@knn_table.register_fake
def _(pos, k, ptr, loop):
    return torch.empty((pos.shape[0], k + (0 if loop else 1)), dtype=torch.int32, device=pos.device)


@_lib.custom_op("dcb200::radius_table", mutates_args=())
def radius_table(pos: Tensor, r: float, ptr: Optional[Tensor], loop: bool, max_num_neighbors: int) -> Tuple[Tensor, Tensor]:
    return ops.radius_table(pos, r, ptr=ptr, loop=loop, max_num_neighbors=max_num_neighbors)


@radius_table.register_fake
def _(pos, r, ptr, loop, max_num_neighbors):
    n = pos.shape[0]
    return (torch.empty((n, max_num_neighbors + (0 if loop else 1)), dtype=torch.int32, device=pos.device),
            torch.empty(n, dtype=torch.int32, device=pos.device))


# ----------------------------------------------------------------------------------- TAGConv
@_lib.custom_op("dcb200::tag_conv", mutates_args=())
def tag_conv(x: Tensor, edge_index: Tensor, weights: List[Tensor], bias: Optional[Tensor], relu: bool, normalize: bool,
             precision: int, ptr: Optional[List[int]]) -> Tuple[Tensor, Tensor]:
    """-> (out [N, Fo], hops [N, K*Fi]) ; hops = [A x | A^2 x | ...] is kept for the backward."""
    g = _csr(edge_index, x.shape[0], "tag" if normalize else "plain", ptr)
    x = x.contiguous()
    N, Fi = x.shape
    K = len(weights) - 1
    buf = torch.empty((N, max(K, 1) * Fi), dtype=x.dtype, device=x.device)
    hs = [g.to_internal(x)]   # a relabelled large graph (ops.REORDER) runs the whole layer in its own node order
    for k in range(K):
        hs.append(buf[:, k * Fi:(k + 1) * Fi])
    if K:
        ops.propagate_chain(g, [(hs[k], None, hs[k + 1]) for k in range(K)], internal=True)   # h_{k+1} = A_hat h_k (one launch, K1 v9)
    out = ops.gemm([(h, w) for h, w in zip(hs, weights)], N, weights[0].shape[0], False, True, bias=bias, relu=relu,
                   precision=precision)
    return g.from_internal(out), buf


@tag_conv.register_fake
def _(x, edge_index, weights, bias, relu, normalize, precision, ptr):
    K = len(weights) - 1
    return x.new_empty((x.shape[0], weights[0].shape[0])), x.new_empty((x.shape[0], max(K, 1) * x.shape[1]))


@_lib.custom_op("dcb200::tag_conv_backward", mutates_args=())
def tag_conv_backward(dout: Tensor, out: Tensor, x: Tensor, hops: Tensor, edge_index: Tensor, weights: List[Tensor],
                      relu: bool, normalize: bool, precision: int, ptr: Optional[List[int]], need_dx: bool, need_db: bool,
                      need_dw: bool) -> Tuple[Tensor, Tensor, Tensor]:
    """-> (dx, dbias, dW stacked [K+1, Fo, Fi]); empty tensors for gradients that are not needed."""
    g = _csr(edge_index, x.shape[0], "tag" if normalize else "plain", ptr)
    dout = dout.contiguous()
    if relu:
        dout = ops.relu_bwd(out, dout)
    dout = g.to_internal(dout)      # the saved hops are in the structure's node order
    N, Fo = dout.shape
    Fi = x.shape[1]
    K = len(weights) - 1
    hs = [g.to_internal(x.contiguous())] + [hops[:, k * Fi:(k + 1) * Fi] for k in range(K)]
    if need_dw:
        dws = torch.empty((K + 1, Fo, Fi), dtype=dout.dtype, device=dout.device)
        for k, h in enumerate(hs):
            ops.gemm([(dout, h)], Fo, Fi, True, False, out=dws[k], precision=precision)
    else:
        dws = dout.new_empty(0)
    db = ops.colsum(dout) if need_db else dout.new_empty(0)
    if need_dx:
        gk = ops.gemm([(dout, weights[K])], N, Fi, False, False, precision=precision)
        if ops.K1_CHAIN >= 2 and K > 0:
            # all dH_k first, then the transposed hops as one chain, accumulating in place: dH_{k-1} += A_hat^T g_k
            dhs = [ops.gemm([(dout, weights[k])], N, Fi, False, False, precision=precision) for k in range(K)] + [gk]
            ops.propagate_chain(g, [(dhs[k + 1], dhs[k], dhs[k]) for k in range(K - 1, -1, -1)], transpose=True, internal=True)
            gk = dhs[0]
        else:
            for k in range(K - 1, -1, -1):
                dhk = ops.gemm([(dout, weights[k])], N, Fi, False, False, precision=precision)
                gk = g.propagate(gk, transpose=True, add=dhk, internal=True)
        dx = g.from_internal(gk)
    else:
        dx = dout.new_empty(0)
    return dx, db, dws


@tag_conv_backward.register_fake
def _(dout, out, x, hops, edge_index, weights, relu, normalize, precision, ptr, need_dx, need_db, need_dw):
    return (torch.empty_like(x) if need_dx else x.new_empty(0), x.new_empty(dout.shape[1]) if need_db else x.new_empty(0),
            x.new_empty((len(weights),) + tuple(weights[0].shape)) if need_dw else x.new_empty(0))


def _tag_setup(ctx, inputs, output):
    x, edge_index, weights, bias, relu, normalize, precision, ptr = inputs
    out, hops = output
    ctx.meta = (relu, normalize, precision, ptr, bias is not None, len(weights))
    ctx.save_for_backward(out, x, hops, edge_index, *weights)
    ctx.set_materialize_grads(False)


def _tag_backward(ctx, dout, dhops):
    relu, normalize, precision, ptr, has_bias, nw = ctx.meta
    out, x, hops, edge_index, *weights = ctx.saved_tensors
    if dout is None:
        return (None,) * 8
    need_dx = ctx.needs_input_grad[0]
    nig_w = ctx.needs_input_grad[2]
    need_dw = any(nig_w) if isinstance(nig_w, (list, tuple)) else bool(nig_w)
    need_db = has_bias and ctx.needs_input_grad[3]
    dx, db, dws = torch.ops.dcb200.tag_conv_backward(dout, out, x, hops, edge_index, list(weights), relu, normalize, precision,
                                                     ptr, need_dx, need_db, need_dw)
    return (dx if need_dx else None, None, list(dws.unbind(0)) if need_dw else None, db if need_db else None, None, None, None,
            None)


tag_conv.register_autograd(_tag_backward, setup_context=_tag_setup)


# ----------------------------------------------------------------------------------- GCNConv
@_lib.custom_op("dcb200::gcn_conv", mutates_args=())
def gcn_conv(x: Tensor, edge_index: Tensor, weight: Tensor, bias: Optional[Tensor], relu: bool, precision: int,
             ptr: Optional[List[int]]) -> Tensor:
    g = _csr(edge_index, x.shape[0], "gcn", ptr)
    xw = ops.gemm([(x.contiguous(), weight)], x.shape[0], weight.shape[0], False, True, precision=precision)
    return g.propagate(xw, bias=bias, relu=relu)


@gcn_conv.register_fake
def _(x, edge_index, weight, bias, relu, precision, ptr):
    return x.new_empty((x.shape[0], weight.shape[0]))


@_lib.custom_op("dcb200::gcn_conv_backward", mutates_args=())
def gcn_conv_backward(dout: Tensor, out: Tensor, x: Tensor, edge_index: Tensor, weight: Tensor, relu: bool, precision: int,
                      ptr: Optional[List[int]], need_dx: bool, need_db: bool) -> Tuple[Tensor, Tensor, Tensor]:
    g = _csr(edge_index, x.shape[0], "gcn", ptr)
    dout = dout.contiguous()
    if relu:
        dout = ops.relu_bwd(out, dout)
    N, Fo = dout.shape
    Fi = x.shape[1]
    db = ops.colsum(dout) if need_db else dout.new_empty(0)
    dxw = g.propagate(dout, transpose=True)
    dx = ops.gemm([(dxw, weight)], N, Fi, False, False, precision=precision) if need_dx else dout.new_empty(0)
    dw = ops.gemm([(dxw, x.contiguous())], Fo, Fi, True, False, precision=precision)
    return dx, dw, db


@gcn_conv_backward.register_fake
def _(dout, out, x, edge_index, weight, relu, precision, ptr, need_dx, need_db):
    return (torch.empty_like(x) if need_dx else x.new_empty(0), torch.empty_like(weight),
            x.new_empty(dout.shape[1]) if need_db else x.new_empty(0))


def _gcn_setup(ctx, inputs, output):
    x, edge_index, weight, bias, relu, precision, ptr = inputs
    ctx.meta = (relu, precision, ptr, bias is not None)
    ctx.save_for_backward(output, x, edge_index, weight)


def _gcn_backward(ctx, dout):
    relu, precision, ptr, has_bias = ctx.meta
    out, x, edge_index, weight = ctx.saved_tensors
    need_dx, need_db = ctx.needs_input_grad[0], has_bias and ctx.needs_input_grad[3]
    dx, dw, db = torch.ops.dcb200.gcn_conv_backward(dout, out, x, edge_index, weight, relu, precision, ptr, need_dx, need_db)
    return (dx if need_dx else None, None, dw, db if need_db else None, None, None, None)


gcn_conv.register_autograd(_gcn_backward, setup_context=_gcn_setup)
