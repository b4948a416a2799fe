for cfg in "4 4" "4 2" "4 1" "8 1" "2 1"; do
  set -- $cfg
  echo "=== DRAIN_KB=$1 SHORT=$2"
  DCB200_DRAIN_KB=$1 DCB200_DRAIN_KB_SHORT=$2 python scripts/acc_lab.py 2>&1 | tail -1
  DCB200_DRAIN_KB=$1 DCB200_DRAIN_KB_SHORT=$2 python scripts/attn_lab.py 4 2>&1 | grep -E "G=4|'gemm', 0" | head -3
  DCB200_DRAIN_KB=$1 DCB200_DRAIN_KB_SHORT=$2 timeout 200 python -m pytest tests/test_gpu_train_loop.py -x -q 2>&1 | grep -E "rel err|passed|failed" | head -3
done
