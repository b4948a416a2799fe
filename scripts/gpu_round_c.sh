#!/bin/bash
# GPU round C (round 2, third session, final code): GPU tests, default bench line + reference arm, full C2 / C4 / C5 lines, labs,
# launch list, full ncu captures of the batched attention products (TMA epilogue), sanitizer
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py 2>gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.json
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_reference.json
timeout 300 python bench.py --workload infer_c2 --steps 10 2>&1 | tail -1 > gpurun_out/bench_c2.json
timeout 300 python bench.py --workload mesh_c4 --steps 5 2>&1 | tail -1 > gpurun_out/bench_c4.json
timeout 300 python bench.py --workload layer_c5 --edges 4096000 --hidden 256 --steps 20 2>&1 | tail -1 > gpurun_out/bench_c5.json
python - <<'PY'
import json
for f in ('bench_default','bench_reference','bench_c2','bench_c4','bench_c5'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['metric'], round(d['value'],2), d.get('e2e') and round(d['e2e']['value'],1), d.get('roofline') and d['roofline'].get('frac') and round(d['roofline']['frac'],3), d.get('roofline_k1') and round(d['roofline_k1']['frac'],3), d.get('gpu_launches'), round(d['ms_per_step'],2), d.get('our_kernel_ms_per_step'))
    except Exception as e:
        print(f, 'FAILED', e)
PY
python scripts/gemm_lab.py > gpurun_out/gemm_lab.txt 2>&1; cat gpurun_out/gemm_lab.txt | cut -c1-120
python scripts/step_calls.py --min-ms 0.3 > gpurun_out/step_calls.txt 2>&1; tail -3 gpurun_out/step_calls.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 90000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --step eager --no-cpu-baseline --no-e2e --no-all-configs > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches.csv
# full captures: the batched attention products of the 4th step (scores Q K^T, P Xr, ..., the fused softmax-backward product)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 188 -c 14 -f -o gpurun_out/prof_gemm_step \
    python bench.py --steps 1 --warmup 3 --step eager --no-cpu-baseline --no-e2e --no-all-configs > gpurun_out/ncu_gemm.log 2>&1
ncu -i gpurun_out/prof_gemm_step.ncu-rep --page raw --csv > gpurun_out/prof_gemm_step_raw.csv 2>/dev/null; wc -l gpurun_out/prof_gemm_step_raw.csv
for t in memcheck racecheck; do timeout 600 compute-sanitizer --tool $t python scripts/sanitize.py > gpurun_out/san_$t.log 2>&1; echo "$t rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize run ok" gpurun_out/san_$t.log | tail -2; done
