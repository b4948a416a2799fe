mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "colsum" 2>&1 | tail -1
for w in 3600 8192; do
  echo "== DCB200_TILE_WHOLE_MAX=$w"
  DCB200_TILE_WHOLE_MAX=$w timeout 300 python bench.py --workload infer_c2 --steps 10 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'graphs/s', round(d['ms_per_step'],2), 'ms  encoder_only', d.get('encoder_only_ms'), d.get('gpu_launches'))"
done
python scripts/step_calls.py --min-ms 1.5 | grep -v "gemm   " | tail -5
python bench.py --no-all-configs --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'graphs/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1), d.get('our_kernel_ms_per_step'))"
