import sys, torch
sys.path.insert(0, ".")
import deformcontact_b200 as dc
from deformcontact_b200 import ops
torch.manual_seed(0)
def run(K, M, N):
    A = torch.randn(K, M, device="cuda"); B = torch.randn(K, N, device="cuda")
    ref = A.double().t() @ B.double()
    out = ops.gemm([(A, B)], M, N, True, False, precision=ops.GEMM_TF32X3)
    torch.cuda.synchronize()
    simt = ops.gemm([(A, B)], M, N, True, False, precision=ops.GEMM_FP32)
    e = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    es = (simt.double() - ref).abs().max().item() / ref.abs().max().item()
    print(f"K={K} M={M} N={N}: rel err tc {e:.3e}  simt {es:.3e}", flush=True)
run(32, 128, 32)
run(1000, 256, 256)
run(333, 128, 32)
run(5000, 200, 64)
run(70000, 256, 256)
run(512000, 256, 256)
K=512000; A=torch.randn(K,256,device="cuda"); B=torch.randn(K,256,device="cuda")
for prec,name in ((ops.GEMM_TF32X3,"tc"),(ops.GEMM_FP32,"simt")):
    for _ in range(3): ops.gemm([(A,B)], 256, 256, True, False, precision=prec)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.gemm([(A,B)], 256, 256, True, False, precision=prec)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/10; print(name, ms, "ms", 2*K*256*256/ms/1e9, "TFLOP/s")
