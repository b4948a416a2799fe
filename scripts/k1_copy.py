"""Tile x slice access-pattern probe: with an edgeless graph and a fused addend, every K1 variant degenerates to
out = add, i.e. a copy that walks memory exactly like the hop kernels (one CTA per tile x 128-byte slice)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deformcontact_b200 import ops
dev = torch.device("cuda", 0)
B, n, F = (int(sys.argv[1]), 2000, int(sys.argv[2])) if len(sys.argv) > 2 else (256, 2000, 256)
N = B * n
ei = torch.zeros(2, 0, dtype=torch.long, device=dev)
G = ops.GraphCSR(ei, N, "tag", [i * n for i in range(B + 1)])
x = torch.randn(N, F, device=dev); a = torch.randn(N, F, device=dev); out = torch.empty_like(x)
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
byt = 2 * N * F * 4
print("torch copy_", round(byt / t(lambda: out.copy_(a)) / 1e6, 1), "GB/s")
for v in ["generic", "lean", "blocks:4"]:
    name, _, fl = v.partition(":")
    ops.K1_VARIANT = name
    if fl: ops.K1_FLAGS = int(fl)
    if name == "generic":
        fn = lambda: ops.spmm(G.rowptr, G.nbr, x, dis=G.dis, add=a, out=out)
    else:
        fn = lambda: G.propagate(x, add=a, out=out)
    fn(); assert torch.equal(out, a)
    ms = t(fn)
    print(v, round(ms, 4), "ms", round(byt / ms / 1e6, 1), "GB/s")
