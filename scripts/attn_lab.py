"""Attention lab: per-op GPU time (CUDA events) and host time of one attention group, fwd + bwd."""
import os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deformcontact_b200 as dc
from deformcontact_b200 import ops, attention
dev = "cuda"
F, ns, nr, G = 256, 8000, 3048, int(sys.argv[1]) if len(sys.argv) > 1 else 4
g = torch.Generator(device=dev).manual_seed(0)
xs = torch.randn(ns * G, F, generator=g, device=dev).relu().requires_grad_(True)
xr = torch.randn(nr * G, F, generator=g, device=dev).relu().requires_grad_(True)
W = ((torch.rand(F, F, generator=g, device=dev) * 2 - 1) / 32).requires_grad_(True)
b = ((torch.rand(F, generator=g, device=dev) * 2 - 1) / 16).requires_grad_(True)
go = torch.randn(ns * G, F, generator=g, device=dev)
groups = [(i * ns, (i + 1) * ns, i * nr, (i + 1) * nr) for i in range(G)]
def step():
    out = attention._AttnFn.apply(xs, xr, groups, W, b)[0]
    out.backward(go)
for _ in range(2): step()
torch.cuda.synchronize()
t0 = time.perf_counter(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); th = time.perf_counter() - t0
torch.cuda.synchronize()
print(f"G={G}: host {th*1e3:.2f} ms, gpu {e0.elapsed_time(e1):.2f} ms per head step; flops {G*6*2*ns*nr*F/1e12:.2f} TF")
ops.PROFILER = []
step(); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in ops.PROFILER:
    k = (r["op"], r.get("M"), r.get("N"), r.get("K"))
    ms = r["e0"].elapsed_time(r["e1"])
    agg[k][0] += 1; agg[k][1] += ms; agg[k][2] += r.get("flops", 0)
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(k, f"n={v[0]} {v[1]:.3f} ms  {v[2]/v[1]/1e9 if v[1] else 0:.1f} TF/s")
ops.PROFILER = None
# torch reference timing
xs2, xr2 = xs.detach().view(G, ns, F).requires_grad_(True), xr.detach().view(G, nr, F).requires_grad_(True)
lin = torch.nn.Linear(F, F).cuda()
def tstep():
    q, k = lin(xs2), lin(xr2)
    o = torch.softmax(q @ k.transpose(-1, -2), -1) @ xr2
    o.backward(go.view(G, ns, F))
for _ in range(2): tstep()
torch.cuda.synchronize(); e0.record(); tstep(); e1.record(); torch.cuda.synchronize()
print(f"torch fp32 head step: {e0.elapsed_time(e1):.2f} ms")
