mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_attention.py -q -x -k "gemm or attention or mlp" 2>&1 | tail -8
for m in 1 0 2; do
  echo "== DCB200_T2_TEPI=$m attn_lab 16"
  DCB200_T2_TEPI=$m timeout 150 python scripts/attn_lab.py 16 2>&1 | head -9
done
for m in 1 0; do
  echo "== DCB200_T2_TEPI=$m gemm_lab"
  DCB200_T2_TEPI=$m timeout 150 python scripts/gemm_lab.py 2>&1 | cut -c1-100
done
