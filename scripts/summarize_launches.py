"""ncu launch list (gpu__time_duration.sum per launch) -> per-kernel summary of the LAST training step.
usage: python scripts/summarize_launches.py gpurun_out/launches.csv profiles/r01_launches_train_c3_v2.csv"""
import csv, re, sys, collections
src, dst = sys.argv[1], sys.argv[2]
rows = []
with open(src) as f:
    rd = csv.reader(l for l in f if l.startswith('"'))
    hdr = next(rd)
    for r in rd:
        if len(r) == len(hdr):
            rows.append(dict(zip(hdr, r)))
rows = [r for r in rows if r["Metric Name"] == "gpu__time_duration.sum"]
ends = [i for i, r in enumerate(rows) if "FusedAdam" in r["Kernel Name"]]   # one fused Adam launch closes every step
last = rows[ends[-2] + 1: ends[-1] + 1] if len(ends) >= 2 else rows
OURS = ("dcb::", "spmm_", "blk_", "rs_", "csr_", "softmax_", "colsum", "relu_bwd", "knn", "pack_edges", "edge_weights", "deg_inv",
        "make_keys", "scan_", "t2_reduce", "gemm_tc2", "sgemm_kernel", "splitk", "mesh_", "posenc", "gat_", "segment_sum", "edge_relu",
        "rowptr", "nbr_", "rowdot", "edge_loss", "batch_vector", "edges_offset", "node_features", "instance_points")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in last:
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    ms = v / 1e6 if u.startswith("n") else v / 1e3 if u.startswith("u") else v
    k = re.sub(r"\(.*", "", r["Kernel Name"])
    k = re.sub(r"<unnamed>::|\(anonymous namespace\)::|void ", "", k)[:110]
    agg[k][0] += 1
    agg[k][1] += ms
tot = sum(v[1] for v in agg.values())
ours = sum(v[1] for k, v in agg.items() if any(o in k for o in OURS))
n_ours = sum(v[0] for k, v in agg.items() if any(o in k for o in OURS))
with open(dst, "w") as f:
    f.write("# ncu launch list, last (4th) training step of `bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e` (C3, 256 graphs/GPU)\n")
    f.write("# command: ncu --metrics gpu__time_duration.sum --clock-control none -c 90000 --csv ...\n")
    f.write("# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes\n")
    f.write(f"# total {tot:.2f} ms over {len(last)} launches; libdcb200 kernels {ours:.2f} ms ({ours / tot:.3f}) in {n_ours} launches\n")
    f.write("kernel,launches,total_ms,share,ours\n")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f'"{k}",{v[0]},{v[1]:.3f},{v[1] / tot:.4f},{int(any(o in k for o in OURS))}\n')
print(open(dst).read()[:2500])
