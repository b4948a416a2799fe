"""Batched kNN-8 build: tiled brute force (dc_knn) against one grid per graph (dc_knn_grid_batched), same output."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deformcontact_b200 as dc
from deformcontact_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
def t(fn, reps=7):
    for _ in range(2): fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
for B, n in ((64, 5000), (256, 2000), (512, 1000), (1024, 500), (2048, 250), (16, 20000)):
    pos = torch.rand(B * n, 3, generator=g, device="cuda")
    ptr = torch.arange(0, B * n + 1, n, device="cuda")
    res = {}
    for mode in ("brute", "grid"):
        ops.KNN_MODE = mode
        res[mode] = (t(lambda: ops.knn_table(pos, 8, ptr=ptr)), ops.knn_table(pos, 8, ptr=ptr))
    ops.KNN_MODE = "auto"
    same = torch.equal(res["brute"][1], res["grid"][1])
    rr = (3.0 * 8 / (4.0 * 3.141592653589793 * n)) ** (1.0 / 3.0)   # ~8 neighbours per point
    rad = {}
    for mode in ("brute", "grid"):
        ops.KNN_MODE = mode
        rad[mode] = (t(lambda: ops.radius_table(pos, rr, ptr=ptr)), ops.radius_table(pos, rr, ptr=ptr)[0])
    ops.KNN_MODE = "auto"
    print(f"{B} x {n}: radius (r for ~8 hits, max 32) brute {rad['brute'][0]:.3f} ms, grid per graph {rad['grid'][0]:.3f} ms, identical {torch.equal(rad['brute'][1], rad['grid'][1])}")
    print(f"{B} x {n}: brute {res['brute'][0]:.3f} ms, grid per graph {res['grid'][0]:.3f} ms (took the grid: {ops.grid_took_it(res['grid'][1])}), identical {same}")
