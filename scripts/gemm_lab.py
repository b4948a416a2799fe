"""GEMM lab: rate and error vs fp64 of dc_gemm (tcgen05 3xTF32) on the shapes the model uses."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deformcontact_b200 as dc
from deformcontact_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
shapes = [  # M, N, K, trans_a, trans_b, label
    (512000, 256, 1024, False, True, "layer fwd"),
    (512000, 1024, 256, False, False, "layer dX"),
    (256, 1024, 512000, True, False, "layer dW"),
    (256, 256, 512000, True, False, "dW 1 hop"),
    (256, 24, 512000, True, False, "dW layer1"),
    (8000, 3048, 256, False, True, "scores"),
    (8000, 256, 3048, False, False, "attn @ Xr"),
    (3048, 256, 8000, True, False, "P^T dO"),
]
def t(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for M, N, K, ta, tb, label in shapes:
    A = torch.randn((K, M) if ta else (M, K), generator=g, device="cuda")
    B = torch.randn((N, K) if tb else (K, N), generator=g, device="cuda")
    out = torch.empty(M, N, device="cuda")
    fn = lambda: ops.gemm([(A, B)], M, N, trans_a=ta, trans_b=tb, out=out, precision=dc._abi.GEMM_TF32X3)
    ms = t(fn)
    # error vs fp64 on a row sample
    rows = torch.randint(0, M, (64,), device="cuda")
    Ar = (A[:, rows].T if ta else A[rows]).double()
    ref = Ar @ (B.double().T if tb else B.double())
    err = ((out[rows].double() - ref).abs().max() / ref.abs().max()).item()
    big = ref.abs() > 0.25 * ref.abs().max()      # signed relative error where |C| is large: < 0 = systematic shrink (RZ accumulate)
    bias = (((out[rows].double() - ref) / ref)[big]).mean().item()
    fn32 = lambda: ops.gemm([(A, B)], M, N, trans_a=ta, trans_b=tb, out=out, precision=dc._abi.GEMM_FP32)
    ms32 = t(fn32, 2)
    err32 = ((out[rows].double() - ref).abs().max() / ref.abs().max()).item()
    print(f"{label:10s} {M}x{N}x{K}: tc {ms:.3f} ms {2.0*M*N*K/ms/1e9:.1f} TF/s err {err:.2e} bias {bias:+.1e} | fp32 simt {ms32:.3f} ms {2.0*M*N*K/ms32/1e9:.1f} TF/s err {err32:.2e}")
