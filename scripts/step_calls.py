"""Every libdcb200 call of one eager C3 train step with its CUDA-event time (ops.PROFILER): which calls carry the step.
usage: python scripts/step_calls.py [--graphs 256] [--min-ms 0.3]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deformcontact_b200 as dc
from deformcontact_b200 import ops, synthetic, model as dcm
dcm.BRANCH_STREAMS = False   # one stream: every call is timed alone

ap = argparse.ArgumentParser()
ap.add_argument("--graphs", type=int, default=256)
ap.add_argument("--nodes", type=int, default=2000)
ap.add_argument("--min-ms", type=float, default=0.3)
a = ap.parse_args()
dev = torch.device("cuda", 0)
rest, rigid, deformed = synthetic.make_batch(a.graphs, a.nodes, 8, first=0, device=dev)
torch.manual_seed(0)
model = dc.load_model(attn_group=4).to(dev)

def step():
    ops.clear_csr_cache()
    model.zero_grad()
    dc.train_step_loss(model, rest, rigid, deformed)[0].backward()

for _ in range(2):
    step()
torch.cuda.synchronize()
prof = []
ops.PROFILER = prof
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record()
torch.cuda.synchronize()
ops.PROFILER = None
tot = 0.0
by = {}
for i, r in enumerate(prof):
    ms = r["e0"].elapsed_time(r["e1"])
    tot += ms
    by[r["op"]] = by.get(r["op"], 0.0) + ms
    if ms >= a.min_ms:
        extra = ""
        if r["op"] == "gemm":
            extra = f"M={r['M']} N={r['N']} K={r['K']}  {r['flops'] / ms / 1e9:7.1f} TFLOP/s"
        elif r["op"] == "spmm":
            extra = f"N={r['N']} F={r['F']} E={r['E']} hops={r.get('hops', 1)}  {r['bytes'] / ms / 1e6:7.1f} GB/s"
        print(f"{i:4d} {r['op']:10s} {ms:8.3f} ms  {extra}")
print(f"step {e0.elapsed_time(e1):.2f} ms wall (eager), {tot:.2f} ms inside {len(prof)} profiled calls; by op: " +
      ", ".join(f"{k} {v:.2f}" for k, v in sorted(by.items(), key=lambda x: -x[1])))
