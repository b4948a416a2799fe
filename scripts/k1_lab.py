"""K1 lab: time the hop kernel variants on the C5 graph (forward and transposed) and check each against the
generic kernel bit for bit.  usage: python scripts/k1_lab.py [--morton] [--nodes 2000] [--graphs 256] [--k 8] [--F 256]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deformcontact_b200 as dc
from deformcontact_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=2000)
ap.add_argument("--graphs", type=int, default=256)
ap.add_argument("--k", type=int, default=8)
ap.add_argument("--F", type=int, default=256)
ap.add_argument("--morton", action="store_true")
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--variants", default="generic,lean,blocks:4,staged:0,staged:1")
ap.add_argument("--graph", default="knn", choices=["knn", "self", "window"])
ap.add_argument("--out", default="")
a = ap.parse_args()
dev = torch.device("cuda", 0)
B, n, k, F = a.graphs, a.nodes, a.k, a.F
N = B * n
g = torch.Generator(device=dev).manual_seed(0)
pos = torch.rand(N, 3, generator=g, device=dev) - 0.5
if a.morton:   # sort the points of every graph along a Morton curve (mesh-like locality)
    q = ((pos + 0.5) * 1023).long().clamp(0, 1023)
    def spread(v):
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v
    code = spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)
    gid = torch.arange(N, device=dev) // n
    order = torch.argsort(gid * (1 << 31) + code)
    pos = pos[order].contiguous()
ptr = torch.arange(B + 1, device=dev) * n
if a.graph == "knn":
    ei = dc.knn_graph(pos, k, ptr=ptr)
else:   # synthetic locality extremes: every edge a self loop / a sliding window of the k next nodes of the graph
    dst = torch.arange(N, device=dev).repeat_interleave(k)
    j = torch.arange(k, device=dev).repeat(N)
    src = dst if a.graph == "self" else (dst // n) * n + (dst % n + j + 1) % n
    ei = torch.stack([src, dst])
E = ei.shape[1]
x = torch.randn(N, F, generator=g, device=dev)
G = ops.GraphCSR(ei, N, "tag", [i * n for i in range(B + 1)])
_ = G.t
ref = {False: ops.spmm(G.rowptr, G.nbr, x, dis=G.dis), True: ops.spmm(G.t[0], G.t[1], x, dis=G.dis)}
out = torch.empty_like(x)
peak = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", 6650.0) if os.path.exists("MEASURED_PEAKS.json") else 6650.0
hop_bytes = 8 * N * F + 4 * E + 8 * N + 4


def ev_time(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


res = []
for v in a.variants.split(","):
    name, _, fl = v.partition(":")
    ops.K1_VARIANT = name
    if fl:
        ops.K1_FLAGS = int(fl)
    row = {"variant": v}
    for tr in (False, True):
        if name == "generic":
            rp, nb = (G.t[0], G.t[1]) if tr else (G.rowptr, G.nbr)
            fn = lambda: ops.spmm(rp, nb, x, dis=G.dis, out=out)
        else:
            fn = lambda: G.propagate(x, transpose=tr, out=out)
        out.zero_()
        fn()
        ok = torch.equal(out, ref[tr])
        ms = ev_time(fn, a.reps)
        row["T" if tr else "fwd"] = {"ms": round(ms, 4), "frac": round(hop_bytes / (ms * 1e-3) / 1e9 / peak, 3), "bit_equal": ok}
    res.append(row)
    print(json.dumps(row), flush=True)
if a.out:
    json.dump({"config": vars(a), "N": N, "E": E, "hop_bytes": hop_bytes, "peak_gbs": peak, "results": res}, open(a.out, "w"), indent=1)
