mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -2 | tee gpurun_out/pytest_gpu.txt
timeout 900 python bench.py 2>gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.json
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_reference.json
timeout 300 python bench.py --workload infer_c2 --steps 10 2>&1 | tail -1 > gpurun_out/bench_c2.json
timeout 300 python bench.py --workload mesh_c4 --steps 5 2>&1 | tail -1 > gpurun_out/bench_c4.json
timeout 300 python bench.py --workload layer_c5 --edges 4096000 --hidden 256 --steps 20 2>&1 | tail -1 > gpurun_out/bench_c5.json
python scripts/step_calls.py --min-ms 0.3 > gpurun_out/step_calls.txt 2>&1; tail -2 gpurun_out/step_calls.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 90000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --step eager --no-cpu-baseline --no-e2e --no-all-configs > gpurun_out/bench_under_ncu.log 2>&1
python - <<'PY'
import json
for f in ('bench_default','bench_reference','bench_c2','bench_c4','bench_c5'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['metric'], round(d['value'],2), d.get('e2e') and round(d['e2e']['value'],1), d.get('roofline') and d['roofline'].get('frac') and round(d['roofline']['frac'],3), d.get('roofline_k1') and round(d['roofline_k1']['frac'],3), d.get('gpu_launches'), round(d['ms_per_step'],2), d.get('our_kernel_ms_per_step'))
    except Exception as e:
        print(f, 'FAILED', e)
PY
