#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "tiled" 2>&1 | tail -2
for v in tiled8 lean; do
  DCB200_K1=$v timeout 300 python bench.py --workload layer_c5 --edges 4096000 --hidden 256 --steps 20 2>&1 | tail -1 > gpurun_out/layer_c5_$v.json
  python - <<PY
import json; d=json.load(open("gpurun_out/layer_c5_$v.json")); print("$v", d["hop"], "frac", round(d["roofline"]["frac"],3), "fb_ms", round(d["ms_per_step"],2))
PY
done
DCB200_K1=lean timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_lean -s 3 -c 1 -f -o gpurun_out/prof_spmm_v6 \
    python bench.py --workload layer_c5 --edges 4096000 --hidden 256 --steps 2 --warmup 3 > gpurun_out/ncu_spmm.log 2>&1
