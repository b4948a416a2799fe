#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_kernels.py tests/test_gpu_layers.py -m gpu -q -x -k "attention or gemm or layer or graphnet or softmax" 2>&1 | tail -15
timeout 300 python scripts/attn_lab.py 4 2>&1 | tail -12
