for d in 4 8 2; do
  echo "== DRAIN_KB=$d attn_lab 16"
  DCB200_DRAIN_KB=$d DCB200_DRAIN_KB_SHORT=$d timeout 150 python scripts/attn_lab.py 16 2>&1 | head -3
  DCB200_DRAIN_KB=$d DCB200_DRAIN_KB_SHORT=$d timeout 150 python scripts/gemm_lab.py 2>&1 | cut -c1-86 | grep -E "layer fwd|layer dX|scores|attn"
done
