for v in 1 0; do
  echo "== DCB200_K1_RECS_SMEM=$v"
  DCB200_K1_RECS_SMEM=$v timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "hop_chain or spmm or hop" 2>&1 | tail -1
  DCB200_K1_RECS_SMEM=$v LAB_SHORT_STREAM=0 timeout 300 python scripts/k1_chain_lab.py 2>&1 | grep -E "fwd_chain |fwd_chain_1hop|fwd_chain_2hop|T_chain |T_bit"
done
