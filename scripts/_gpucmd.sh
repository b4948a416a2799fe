mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 90000 --csv --log-file gpurun_out/launches_b32.csv \
    python bench.py --global-batch 32 --steps 1 --warmup 3 --step eager --no-cpu-baseline --no-e2e --no-all-configs > gpurun_out/bench_under_ncu_b32.log 2>&1
wc -l gpurun_out/launches_b32.csv
python bench.py --global-batch 32 --steps 30 --no-all-configs --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('graph replay, 32 graphs:', round(d['ms_per_step'],3), 'ms/step', d.get('gpu_launches_per_step'), 'eager', d.get('eager_step_ms'))"
