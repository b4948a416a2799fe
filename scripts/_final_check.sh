mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py 2>gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.json
timeout 300 python bench.py --workload infer_c2 --steps 10 2>&1 | tail -1 > gpurun_out/bench_c2.json
python - <<'PY'
import json
for f in ('bench_default','bench_c2'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['metric'], round(d['value'],2), d.get('e2e') and round(d['e2e']['value'],1), d.get('roofline') and d['roofline'].get('frac') and round(d['roofline']['frac'],3), d.get('roofline_k1') and round(d['roofline_k1']['frac'],3), d.get('gpu_launches'), round(d['ms_per_step'],2), d.get('our_kernel_ms_per_step'), d.get('knn_build_ms'), d['clocks'])
    except Exception as e:
        print(f, 'FAILED', e)
PY
