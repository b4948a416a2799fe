# round 2 session 2, call B: stream-kernel parity + timing, fused-epilogue GEMM timing, whole GPU suite, default bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "hop_chain" 2>&1 | tail -3
timeout 300 python scripts/k1_chain_lab.py --out gpurun_out/k1_chain_lab_v11.json 2>&1 | tail -16
DCB200_K1_STREAM_THREADS=1024 timeout 300 python scripts/k1_chain_lab.py 2>&1 | grep -E "v11|T_bit"
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
timeout 900 python bench.py --no-all-configs 2>gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default_b.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default_b.json')); print(d['metric'], round(d['value'],2), d.get('e2e') and round(d['e2e']['value'],1), round(d['roofline']['frac'],3), round(d['roofline_k1']['frac'],3), d.get('gpu_launches'), round(d['ms_per_step'],2))
print(d.get('our_kernel_ms_per_step'))
PY
