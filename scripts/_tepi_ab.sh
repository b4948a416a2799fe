mkdir -p gpurun_out
echo skip tests
for cfg in "0 0" "1 0" "1 1"; do
  set -- $cfg
  echo "== TEPI=$1 HINT=$2"
  DCB200_T2_TEPI=$1 DCB200_T2_TEPI_HINT=$2 timeout 300 python bench.py --no-all-configs --no-cpu-baseline --steps 10 --warmup 3 2>gpurun_out/ab.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],2), d.get('e2e',{}).get('value'), d.get('our_kernel_ms_per_step'), d['clocks'])"
done
