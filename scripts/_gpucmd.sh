mkdir -p gpurun_out
echo "== tests with PAIR=1 (default)"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for pr in 1 0 1 0; do echo "== PAIR=$pr"; DCB200_T2_PAIR=$pr timeout 300 python scripts/step_calls.py --min-ms 7 | tail -7; done
