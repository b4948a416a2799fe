#!/bin/bash
# GPU round A (round 2): GPU tests, default bench line + reference arm, labs, launch list, full ncu captures of K2 and K1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
timeout 900 python bench.py 2>gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.json
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_reference.json
python - <<'PY'
import json
for f in ('bench_default','bench_reference'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['metric'], round(d['value'],2), d.get('e2e') and round(d['e2e']['value'],1), d.get('roofline') and d['roofline'].get('frac') and round(d['roofline']['frac'],3), d.get('roofline_k1') and round(d['roofline_k1']['frac'],3), d.get('gpu_launches'), round(d['ms_per_step'],2))
        for k in ('c2','c4','c5'):
            if k in d: print('  ', k, d[k])
    except Exception as e:
        print(f, 'FAILED', e)
PY
python scripts/gemm_lab.py > gpurun_out/gemm_lab.txt 2>&1; cat gpurun_out/gemm_lab.txt
timeout 300 python scripts/k1_chain_lab.py --out gpurun_out/k1_chain_lab.json > gpurun_out/k1_chain_lab.txt 2>&1; tail -20 gpurun_out/k1_chain_lab.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 90000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --step eager --no-cpu-baseline --no-e2e --no-all-configs > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 100 -c 3 -f -o gpurun_out/prof_gemm_step \
    python bench.py --steps 1 --warmup 3 --step eager --no-cpu-baseline --no-e2e --no-all-configs > gpurun_out/ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_chain_kernel -s 6 -c 1 -f -o gpurun_out/prof_spmm_c5 \
    python bench.py --workload layer_c5 --edges 4096000 --hidden 256 --steps 3 > gpurun_out/ncu_spmm_c5.log 2>&1
python scripts/ncu_traffic.py gpurun_out/r02_ncu_traffic.json "gemm_tc2_kernel=gpurun_out/prof_gemm_step.ncu-rep:bench.py train step, launches 101-103 of gemm_tc2_kernel" \
    "spmm_chain_kernel_c5=gpurun_out/prof_spmm_c5.ncu-rep:bench.py --workload layer_c5, 3-hop forward chain N=512000 F=256 k=8"
ls -la gpurun_out
