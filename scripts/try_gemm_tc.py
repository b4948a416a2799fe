import sys, torch
sys.path.insert(0, ".")
import deformcontact_b200 as dc
from deformcontact_b200 import ops
torch.manual_seed(0)
def run(M, N, Ks, tb=True, bias=True, relu=False):
    As = [torch.randn(M, k, device="cuda") for k in Ks]
    Bs = [torch.randn(N, k, device="cuda") if tb else torch.randn(k, N, device="cuda") for k in Ks]
    b = torch.randn(N, device="cuda") if bias else None
    ref = sum(a.double() @ (w.double().t() if tb else w.double()) for a, w in zip(As, Bs))
    if bias: ref = ref + b.double()
    if relu: ref = ref.clamp_min(0)
    out = ops.gemm(list(zip(As, Bs)), M, N, False, tb, bias=b, relu=relu, precision=ops.GEMM_TF32X3)
    torch.cuda.synchronize()
    simt = ops.gemm(list(zip(As, Bs)), M, N, False, tb, bias=b, relu=relu, precision=ops.GEMM_FP32)
    e = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    es = (simt.double() - ref).abs().max().item() / ref.abs().max().item()
    print(f"M={M} N={N} K={Ks} tb={tb}: rel err tc {e:.3e}  simt {es:.3e}", flush=True)
    return e
run(128, 256, [32])
run(128, 256, [256])
run(300, 64, [32, 64])
run(1000, 256, [256, 256, 256, 256], relu=True)
run(5000, 128, [96], tb=False)
run(70000, 256, [256] * 4)
run(512000, 256, [256] * 4)
import time
M=512000; As=[torch.randn(M,256,device="cuda") for _ in range(4)]; Bs=[torch.randn(256,256,device="cuda") for _ in range(4)]
for prec,name in ((ops.GEMM_TF32X3,"tc"),(ops.GEMM_FP32,"simt")):
    for _ in range(3): ops.gemm(list(zip(As,Bs)), M, 256, False, True, precision=prec)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.gemm(list(zip(As,Bs)), M, 256, False, True, precision=prec)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/10; print(name, ms, "ms", 2*M*256*1024/ms/1e9, "TFLOP/s")
