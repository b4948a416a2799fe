#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --workload layer_c5 --edges 4096000 --hidden 256 --steps 20 2>&1 | tail -1 | tee gpurun_out/layer_c5.json
timeout 600 python bench.py --workload layer_c5 --edges 4096000 --hidden 64 --steps 20 2>&1 | tail -1 | tee gpurun_out/layer_c5_f64.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_tiled_kernel -s 3 -c 1 -f -o gpurun_out/prof_spmm_tiled_fwd \
    python bench.py --workload layer_c5 --edges 4096000 --hidden 256 --steps 2 --warmup 3 > gpurun_out/ncu_spmm.log 2>&1
