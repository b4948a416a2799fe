# driver-like final check: GPU tests, smoke(), default bench line, reference arm
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py 2>gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.json
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_reference.json
python - <<'PY'
import json
for f in ('bench_default','bench_reference'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['metric'], round(d['value'],2), d.get('e2e') and round(d['e2e']['value'],1), d.get('roofline') and d['roofline'].get('frac') and round(d['roofline']['frac'],3), d.get('roofline_k1') and round(d['roofline_k1']['frac'],3), d.get('gpu_launches'), round(d['ms_per_step'],2), d.get('our_kernel_ms_per_step'), d['clocks'] if 'clocks' in d else '')
    if 'roofline' in d: print('  traffic', d['roofline'].get('traffic'), d['roofline_k1'].get('traffic'))
    for k in ('c2','c4','c5'):
        if k in d: print('  ', k, round(d[k]['value'],2), round(d[k]['ms_per_step'],2))
PY
