"""K1 v9 / v10 lab (v10 = dc_spmm_stage, tile slice staged in shared memory by TMA): the three hops of a TAGConv layer as one chain launch vs three launches, forward (lean) and
transposed-with-addend (blocks / lean), bit-equality against the per-hop kernels, C5 graph.
usage: python scripts/k1_chain_lab.py [--nodes 2000] [--graphs 256] [--k 8] [--F 256] [--out file.json]"""
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
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--out", default="")
a = ap.parse_args()
dev = torch.device("cuda", 0)
B, n, k, F = a.graphs, a.nodes, a.k, a.F
N = B * n
gen = torch.Generator(device=dev).manual_seed(0)
pos = torch.rand(N, 3, generator=gen, device=dev) - 0.5
ptr = torch.arange(B + 1, device=dev) * n
ei = dc.knn_graph(pos, k, ptr=ptr)
E = ei.shape[1]
G = ops.GraphCSR(ei, N, "tag", [i * n for i in range(B + 1)])
_ = G.t
assert G.tiles_closed
x = torch.randn(N, F, generator=gen, device=dev)
adds = [torch.randn(N, F, generator=gen, device=dev) for _ in range(3)]
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


res = {}
# forward: h1 = A x, h2 = A h1, h3 = A h2 into one [N, 3F] buffer (the TAGConv layout)
buf_a, buf_b = (torch.empty(N, 3 * F, device=dev) for _ in range(2))
views = lambda b: [b[:, i * F:(i + 1) * F] for i in range(3)]
def fwd_sep(b):
    v = views(b); src = x
    for i in range(3):
        ops.spmm_lean(G.rowptr, G.edges, None, src, out=v[i], tile_ptr=G.tile_ptr, n_tiles=G.n_tiles); src = v[i]
def fwd_chain(b, staged=False):
    v = views(b)
    ops.spmm_chain(G.rowptr, G.edges, None, [(x, None, v[0]), (v[0], None, v[1]), (v[1], None, v[2])], tile_ptr=G.tile_ptr, n_tiles=G.n_tiles,
                   max_tile_rows=G._max_tile if staged else 0)
fwd_sep(buf_a); buf_b.zero_(); fwd_chain(buf_b)
ok = torch.equal(buf_a, buf_b)
buf_c = torch.zeros(N, 3 * F, device=dev)
fwd_chain(buf_c, True)
ok_staged = torch.equal(buf_a, buf_c)
ms = ev_time(lambda: fwd_chain(buf_c, True), a.reps)
res["fwd_chain_staged_v10"] = {"ms": round(ms, 4), "ms_per_hop": round(ms / 3, 4), "frac": round(3 * hop_bytes / (ms * 1e-3) / 1e9 / peak, 3), "bit_equal": ok_staged}
print("fwd_chain_staged_v10", json.dumps(res["fwd_chain_staged_v10"]), flush=True)
del buf_c
def with_stream(flag, fn):
    def run():
        ops.K1_STREAM = flag
        fn()
    return run
buf_d = torch.zeros(N, 3 * F, device=dev)
with_stream(1, lambda: fwd_chain(buf_d))()
ok_stream = torch.equal(buf_a, buf_d)
del buf_d
ops.K1_STREAM = 0
for name, fn in (("fwd_3_launches", lambda: fwd_sep(buf_a)), ("fwd_chain", with_stream(0, lambda: fwd_chain(buf_b))),
                 ("fwd_chain_stream_v11", with_stream(1, lambda: fwd_chain(buf_b)))):
    if name.endswith("v11"):
        ok = ok_stream
    ms = ev_time(fn, a.reps)
    res[name] = {"ms": round(ms, 4), "ms_per_hop": round(ms / 3, 4), "frac": round(3 * hop_bytes / (ms * 1e-3) / 1e9 / peak, 3), "bit_equal": ok}
    print(name, json.dumps(res[name]), flush=True)
for hops in (1, 2):
    v = views(buf_b)
    ops.K1_STREAM = int(os.environ.get("LAB_SHORT_STREAM", "1"))
    fn = lambda: ops.spmm_chain(G.rowptr, G.edges, None, [(x, None, v[0]), (v[0], None, v[1])][:hops], tile_ptr=G.tile_ptr, n_tiles=G.n_tiles)
    ms = ev_time(fn, a.reps)
    res[f"fwd_chain_{hops}hop"] = {"ms": round(ms, 4), "ms_per_hop": round(ms / hops, 4), "frac": round(hops * hop_bytes / (ms * 1e-3) / 1e9 / peak, 3)}
    print(f"fwd_chain_{hops}hop", json.dumps(res[f"fwd_chain_{hops}hop"]), flush=True)

# transposed with addends, in place: d2 += A^T g3; d1 += A^T d2; d0 += A^T d1
g3 = torch.randn(N, F, generator=gen, device=dev)
hb = hop_bytes + 4 * N * F
def bwd_sep(variant):
    ops.K1_VARIANT = variant
    d = [t.clone() for t in adds]; src = g3
    def run():
        s = src
        for i in (2, 1, 0):
            s = G.propagate(s, transpose=True, add=d[i], out=d[i])
    return d, run
def bwd_chain(staged=False):
    d = [t.clone() for t in adds]
    def run():
        ops.spmm_chain(G.t[0], G._edges_t, None, [(g3, d[2], d[2]), (d[2], d[1], d[1]), (d[1], d[0], d[0])], tile_ptr=G.tile_ptr, n_tiles=G.n_tiles,
                       max_tile_rows=G._max_tile if staged else 0)
    return d, run
outs = {}
for name, (d, run) in (("T_3_blocks", bwd_sep("blocks")), ("T_3_lean", bwd_sep("lean")), ("T_chain", bwd_chain()), ("T_chain_staged_v10", bwd_chain(True)),
                       ("T_chain_stream_v11", bwd_chain())):
    ops.K1_STREAM = 1 if name.endswith("v11") else 0
    if not name.startswith("T_chain"):
        ops.K1_VARIANT = name.split("_")[-1]
    run(); torch.cuda.synchronize()
    outs[name] = [t.clone() for t in d]
    ms = ev_time(run, a.reps)   # (values drift as the in-place accumulation repeats; timing only)
    res[name] = {"ms": round(ms, 4), "ms_per_hop": round(ms / 3, 4), "frac": round(3 * hb / (ms * 1e-3) / 1e9 / peak, 3)}
    print(name, json.dumps(res[name]), flush=True)
ops.K1_VARIANT = "auto"
eq = (all(torch.equal(p, q) for p, q in zip(outs["T_3_blocks"], outs["T_chain"])) and all(torch.equal(p, q) for p, q in zip(outs["T_3_lean"], outs["T_chain"]))
      and all(torch.equal(p, q) for p, q in zip(outs["T_chain_staged_v10"], outs["T_chain"]))
      and all(torch.equal(p, q) for p, q in zip(outs["T_chain_stream_v11"], outs["T_chain"])))
ops.K1_STREAM = 1
res["T_bit_equal"] = eq
print("T_bit_equal", eq)
if a.out:
    json.dump({"config": vars(a), "N": N, "E": E, "hop_bytes": hop_bytes, "peak_gbs": peak, "results": res}, open(a.out, "w"), indent=1)
