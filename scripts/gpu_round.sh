#!/bin/bash
# one GPU round: GPU tests, the bench lines of every workload, the launch list and the full ncu captures
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.txt
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_default.json
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_reference.json
timeout 300 python bench.py --workload infer_c2 --steps 10 2>&1 | tail -1 > gpurun_out/bench_c2.json
timeout 300 python bench.py --workload mesh_c4 --steps 5 2>&1 | tail -1 > gpurun_out/bench_c4.json
timeout 300 python bench.py --workload layer_c5 --edges 4096000 --hidden 256 --steps 20 2>&1 | tail -1 > gpurun_out/bench_c5.json
python -c "
import json
for f in ('bench_default','bench_reference','bench_c2','bench_c4','bench_c5'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['metric'], round(d['value'],2), d.get('e2e') and round(d['e2e']['value'],1), d.get('roofline') and round(d['roofline']['frac'],3), d.get('gpu_launches'), round(d['ms_per_step'],2))
"
python scripts/gemm_lab.py > gpurun_out/gemm_lab.txt 2>&1
timeout 300 python scripts/k1_chain_lab.py --out gpurun_out/k1_chain_lab.json > gpurun_out/k1_chain_lab.txt 2>&1; cat gpurun_out/k1_chain_lab.txt
python scripts/step_profile.py > gpurun_out/step_profile.txt 2>&1; tail -30 gpurun_out/step_profile.txt | head -14
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 90000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches.csv
for t in memcheck racecheck synccheck; do timeout 400 compute-sanitizer --tool $t python scripts/sanitize.py > gpurun_out/san_$t.log 2>&1; echo "$t rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize run ok" gpurun_out/san_$t.log | tail -2; done
