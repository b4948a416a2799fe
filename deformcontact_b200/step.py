"""The reference's optimisation step (train.py:35-73) as ONE replayable CUDA graph per input shape.

``train.py`` issues, per mini-batch: forward (train.py:46), displacement targets (:47-50), L1 + lambda * consistency loss
(:52-58), ``zero_grad / backward / step`` (:71-73) — on our path ~430 kernel launches through ctypes and the autograd engine.
At BASELINE config C3 split over 8 GPUs (32 graphs per GPU) the kernels of a step take ~12 ms while the host needs longer
than that to issue them, so the step is captured once and replayed:

    structure build (CSR pair per graph batch) -> [batch assembly N3] -> forward -> losses -> backward
        -> gradient all-reduce (data parallel) -> Adam

Everything inside is static per shape: inputs live in fixed device buffers (``run`` copies new data into them), the tile
tables of K1 are device-cached (ops._device_table), the problem tables of ``dc_gemm_batched`` are staged through pinned
buffers this object keeps alive, and the optimizer runs in its capturable form.  Replays are bit-identical to the eager
step (tests/test_gpu_train_loop.py, tests/test_gpu_step.py).

Data parallel: by default the step is two graphs (forward + backward | Adam) around an eager NCCL all-reduce of the flat
gradient buffer; ``DCB200_GRAPH_ALLREDUCE=1`` captures the collective into one graph with the rest (NCCL supports stream
capture; falls back to the two-graph form if the capture fails).
"""
import os

import torch

from . import _abi, ops
from .model import fused_losses

# 1: capture the NCCL all-reduce into the step's graph; 0 (default): two graphs around an eager all-reduce — three host calls
# per step instead of one (~20 us), and no dependence on NCCL's stream-capture support at a rank count that was not tested
GRAPH_ALL_REDUCE = os.environ.get("DCB200_GRAPH_ALLREDUCE", "0") != "0"


def _clone_batch(b):
    """Static copy of a batch: fresh device tensors with the same contents, host-side graph offsets carried over."""
    from .model import _host_ptr
    if getattr(b, "ptr", None) is not None:
        _host_ptr(b)            # reads ptr back once, outside any capture
    out = b.clone()
    for k in ("_ptr_host", "_edge_ptr"):
        if hasattr(b, k):
            setattr(out, k, getattr(b, k))
    return out


def _snapshot_optimizer(opt):
    snap = {}
    for p, st in opt.state.items():
        snap[p] = {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in st.items()}
    return snap


def _restore_optimizer(opt, snap):
    """Put the optimizer state back IN PLACE (state tensors keep their addresses: they are about to be captured).
    State created by the warm-up steps of a fresh optimizer is zeroed, which is exactly its initial value."""
    for p, st in opt.state.items():
        old = snap.get(p)
        for k, v in st.items():
            if torch.is_tensor(v):
                if old is not None and k in old:
                    v.copy_(old[k])
                else:
                    v.zero_()
            elif old is not None and k in old:
                st[k] = old[k]


class CapturedTrainStep:
    """One optimisation step of ``model`` captured in a CUDA graph.

    Two input forms:
      * batches: ``CapturedTrainStep(model, opt, rest, rigid, deformed)`` then ``run(rest, rigid, deformed)`` with batches
        of the same shapes and graph offsets (``x``, ``pos``, ``edge_index`` are copied into the static buffers);
      * raw: ``CapturedTrainStep(model, opt, raw=dict_of_device_tensors, assemble=fn)`` where ``fn(raw) -> (rest, rigid,
        deformed)`` runs INSIDE the graph (batch assembly N3); ``run(raw=dict)`` copies host (pinned) or device tensors into
        the static raw buffers.

    ``flat_grads`` (dist.FlatGrads) + ``loss_shares`` (dist.loss_shares) give the data-parallel step: local losses are
    scaled by this rank's share of nodes / edges and the flat gradient is SUM-all-reduced before Adam.
    ``run`` returns the step's (loss, l1, consistency) as device scalars that the next ``run`` overwrites.
    """

    def __init__(self, model, optimizer, rest=None, rigid=None, deformed=None, lambda_gradient=1.0, raw=None, assemble=None,
                 flat_grads=None, loss_shares=(1.0, 1.0), warmup=2, optimizer_step=True):
        if not torch.cuda.is_available():
            raise _abi.DcError("CapturedTrainStep needs a CUDA device (libdcb200 has no CPU path)")
        for grp in optimizer.param_groups:
            if optimizer_step and grp.get("capturable", True) is False:
                raise _abi.DcError("CapturedTrainStep: construct the optimizer with capturable=True (e.g. torch.optim.Adam(..., "
                                   "capturable=True)); its step is captured into the CUDA graph")
        self.model, self.opt, self.lam = model, optimizer, float(lambda_gradient)
        self.flat, self.shares = flat_grads, (float(loss_shares[0]), float(loss_shares[1]))
        self.optimizer_step = optimizer_step
        self._keep = []          # pinned staging buffers the graph's copy nodes re-read on every replay
        self.graphs = []
        self.launches_per_step = 0
        if raw is not None:
            if assemble is None:
                raise ValueError("raw inputs need an assemble(raw) function")
            self.raw = {k: v.clone() for k, v in raw.items()}
            self._assemble = assemble
            self.static = None
        else:
            self.raw, self._assemble = None, None
            self.static = tuple(_clone_batch(b) for b in (rest, rigid, deformed))
        self._dist = torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1
        self._params = [p for p in model.parameters() if p.requires_grad]
        self._capture(int(warmup))

    # ------------------------------------------------------------------ the step body (eager during warm-up, then captured)
    def _batches(self):
        return self._assemble(self.raw) if self.raw is not None else self.static

    def _zero_grads(self):
        if self.flat is not None:
            self.flat.zero_()
        else:
            for p in self._params:
                if p.grad is not None:
                    p.grad.zero_()

    def _fwd_bwd(self):
        ops.clear_csr_cache()
        rest, rigid, deformed = self._batches()
        self._zero_grads()
        pred = self.model(rest, rigid)
        pred.pos = pred.pos - rest.pos                      # train.py:47-50: displacement fields
        tgt = deformed.clone()
        tgt.pos = deformed.pos - rest.pos
        l1, lc = fused_losses(pred, tgt)                     # train.py:52-56
        loss = self.shares[0] * l1 + (self.lam * self.shares[1]) * lc     # train.py:58 (x this rank's share under DP)
        loss.backward()
        return loss.detach(), l1.detach(), lc.detach()

    def _all_reduce(self):
        if self._dist:
            if self.flat is None:
                raise _abi.DcError("data-parallel CapturedTrainStep needs flat_grads (dist.FlatGrads)")
            torch.distributed.all_reduce(self.flat.flat, op=torch.distributed.ReduceOp.SUM)

    def _eager_step(self):
        out = self._fwd_bwd()
        self._all_reduce()
        if self.optimizer_step:
            self.opt.step()
        return out

    # ------------------------------------------------------------------ capture
    def _capture(self, warmup):
        lib = _abi.lib()
        params_before = [p.detach().clone() for p in self._params]
        opt_before = _snapshot_optimizer(self.opt)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                        # warm-up on a side stream (torch's capture recipe): fills the
            for _ in range(max(warmup, 1)):                  # device table cache, creates .grad and optimizer state, NCCL comm
                self._eager_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.no_grad():                                # the warm-up steps must leave no trace
            for p, q in zip(self._params, params_before):
                p.copy_(q)
            _restore_optimizer(self.opt, opt_before)
        torch.cuda.synchronize()

        ops.CAPTURE_KEEPALIVE = self._keep
        try:
            whole = GRAPH_ALL_REDUCE or not self._dist
            if whole:
                try:
                    g = torch.cuda.CUDAGraph()
                    l0 = lib.dc_launch_count()
                    with torch.cuda.graph(g):
                        self.out = self._eager_step()
                    self.launches_per_step = int(lib.dc_launch_count() - l0)
                    self.graphs = [g]
                    self.mode = "one graph (structure build + forward + loss + backward" + (" + all-reduce" if self._dist else "") + " + Adam)"
                except Exception as e:   # NCCL refused the capture: keep the collective outside
                    if not self._dist:
                        raise
                    self._capture_error = repr(e)
                    torch.cuda.synchronize()
                    whole = False
            if not whole:
                ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                l0 = lib.dc_launch_count()
                with torch.cuda.graph(ga):
                    self.out = self._fwd_bwd()
                self.launches_per_step = int(lib.dc_launch_count() - l0)
                if self.optimizer_step:
                    with torch.cuda.graph(gb):
                        self.opt.step()
                    self.graphs = [ga, gb]
                else:
                    self.graphs = [ga]
                self.mode = "two graphs (forward + backward | Adam) around an eager all-reduce"
        finally:
            ops.CAPTURE_KEEPALIVE = None
            ops.clear_csr_cache()      # structures built during capture live in the graph's pool; do not hand them out
        torch.cuda.synchronize()
        with torch.no_grad():          # capture executes nothing, but be explicit: weights and state as before
            for p, q in zip(self._params, params_before):
                p.copy_(q)
            _restore_optimizer(self.opt, opt_before)

    # ------------------------------------------------------------------ replay
    @staticmethod
    def _check_same(a, b, what):
        if tuple(a.shape) != tuple(b.shape) or a.dtype != b.dtype:
            raise _abi.DcError(f"CapturedTrainStep.run: {what} is {tuple(b.shape)} {b.dtype}, captured {tuple(a.shape)} {a.dtype} "
                               "(capture one step per shape bucket)")

    def load(self, rest=None, rigid=None, deformed=None, raw=None):
        """Copy one step's inputs into the static buffers (asynchronous on the current stream)."""
        if self.raw is not None:
            for k, v in raw.items():
                self._check_same(self.raw[k], v, f"raw['{k}']")
                self.raw[k].copy_(v, non_blocking=True)
            return
        for s, b, name in zip(self.static, (rest, rigid, deformed), ("rest", "rigid", "deformed")):
            if getattr(s, "_ptr_host", None) is not None and getattr(b, "ptr", None) is not None:
                from .model import _host_ptr
                if list(_host_ptr(b)) != list(s._ptr_host):
                    raise _abi.DcError(f"CapturedTrainStep.run: graph offsets of '{name}' differ from the captured batch")
            for k in ("x", "pos", "edge_index"):
                t = getattr(b, k, None)
                if t is not None:
                    self._check_same(getattr(s, k), t, f"{name}.{k}")
                    getattr(s, k).copy_(t, non_blocking=True)

    def replay(self):
        if len(self.graphs) == 1:
            self.graphs[0].replay()
        else:
            self.graphs[0].replay()
            self._all_reduce()
            self.graphs[1].replay()
        return self.out

    def run(self, rest=None, rigid=None, deformed=None, raw=None):
        self.load(rest, rigid, deformed, raw)
        return self.replay()
