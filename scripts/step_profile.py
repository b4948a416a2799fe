"""Kernel-time breakdown of one C3 training step with torch.profiler (CUPTI): which kernels the step spends its time in."""
import os, sys, collections, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deformcontact_b200 as dc
from deformcontact_b200 import synthetic, ops
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
Bg = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.backends.cuda.matmul.allow_tf32 = False
rest, rigid, deformed = synthetic.make_batch(Bg, 2000, 8, first=0, device=dev)
torch.manual_seed(0)
model = dc.load_model(attn_group=4).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=4e-4, fused=True)
def step():
    ops.clear_csr_cache()
    opt.zero_grad(set_to_none=True)
    loss, _, _ = dc.train_step_loss(model, rest, rigid, deformed)
    loss.backward()
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        k = re.sub(r"<.*", "", e.name)[:70]
        agg[k][0] += 1; agg[k][1] += e.device_time / 1e3
tot = sum(v[1] for v in agg.values())
print(f"total kernel time {tot:.2f} ms")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:28]:
    print(f"{v[1]:8.2f} ms {100*v[1]/tot:5.1f}% n={v[0]:5d} {k}")
