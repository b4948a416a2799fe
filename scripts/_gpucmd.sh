mkdir -p gpurun_out
timeout 900 python bench.py 2>gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.json
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_reference.json
python scripts/step_calls.py --min-ms 0.3 > gpurun_out/step_calls.txt 2>&1; tail -1 gpurun_out/step_calls.txt
python - <<'PY'
import json
for f in ('bench_default','bench_reference'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, round(d['value'],2), d.get('e2e') and round(d['e2e']['value'],1), d.get('roofline') and round(d['roofline']['frac'],3), d.get('roofline_k1') and round(d['roofline_k1']['frac'],3), d.get('gpu_launches'), round(d['ms_per_step'],2), d.get('our_kernel_ms_per_step'))
    for k in ('c2','c4','c5'):
        if k in d: print('  ', k, round(d[k]['value'],2), round(d[k]['ms_per_step'],2))
PY
