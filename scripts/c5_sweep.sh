#!/bin/bash
# C5: message-passing layer microbench sweep (hidden 64/128/256, 1M-50M edges, fwd+bwd vs roofline)
mkdir -p gpurun_out
: > gpurun_out/c5_sweep.jsonl
for layer in tag mpnn; do
for F in 64 128 256; do
for E in 1e6 5e6 20e6 50e6; do
  timeout 300 python bench.py --workload layer_c5 --layer $layer --edges $E --hidden $F --steps 5 2>&1 | tail -1 >> gpurun_out/c5_sweep.jsonl
done; done; done
python - <<'PY'
import json
print("layer F E fwd+bwd_ms edges/s hop_ms hop_frac")
for l in open("gpurun_out/c5_sweep.jsonl"):
    try: d=json.loads(l)
    except Exception: print("bad line", l[:100]); continue
    w=d["config"]["workload"]
    print(w.split(" layer")[0][3:], w.split("E=")[1], round(d["ms_per_step"],3), f'{d["value"]:.3e}', round(d["hop"]["fwd_ms"],4), round(d["roofline"]["frac"],3))
PY
