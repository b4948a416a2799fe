for m in 1 3 1 3; do
  echo "== TEPI=$m"
  DCB200_T2_TEPI=$m timeout 300 python bench.py --no-all-configs --no-cpu-baseline --no-e2e --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],2), d.get('our_kernel_ms_per_step'), d['clocks']['sm_mhz'])"
done
