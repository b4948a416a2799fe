#!/bin/bash
# C5: message-passing layer microbench sweep (hidden 64/128/256, 1M-50M edges, k = 8 / 16, node order random / Morton,
# fwd+bwd vs roofline) -> gpurun_out/c5_sweep.jsonl
mkdir -p gpurun_out
: > gpurun_out/c5_sweep.jsonl
for k in 8 16; do
for F in 64 128 256; do
for E in 1e6 5e6 20e6 50e6; do
  timeout 300 python bench.py --workload layer_c5 --layer tag --k $k --edges $E --hidden $F --steps 5 2>&1 | tail -1 >> gpurun_out/c5_sweep.jsonl
done; done; done
for F in 64 256; do
  timeout 300 python bench.py --workload layer_c5 --layer tag --k 8 --order morton --edges 20e6 --hidden $F --steps 5 2>&1 | tail -1 >> gpurun_out/c5_sweep.jsonl
  timeout 300 python bench.py --workload layer_c5 --layer mpnn --k 8 --edges 20e6 --hidden $F --steps 5 2>&1 | tail -1 >> gpurun_out/c5_sweep.jsonl
done
python - <<'PY'
import json
print("layer k order F E fwd+bwd_ms edges/s chain_ms_per_hop hop_frac")
for l in open("gpurun_out/c5_sweep.jsonl"):
    try: d=json.loads(l)
    except Exception: print("bad line", l[:100]); continue
    w=d["config"]["workload"]
    print(w.split(" layer")[0][3:], w.split("kNN-")[1].split(" ")[0], d["config"].get("node_order"), w.split("E=")[1], round(d["ms_per_step"],3),
          f'{d["value"]:.3e}', round(d["hop"]["chain3_ms_per_hop"],4), round(d["roofline"]["frac"],3))
PY
