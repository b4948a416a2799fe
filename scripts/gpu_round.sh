#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.txt
python scripts/gemm_lab.py 2>&1 | tail -8
timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_default.json
python -c "
import json
d=json.load(open('gpurun_out/bench_default.json')); print(d['metric'], round(d['value'],2), d.get('e2e') and round(d['e2e']['value'],1), d.get('gpu_launches'), round(d['ms_per_step'],2), d['roofline_tensor']['achieved'])
"
