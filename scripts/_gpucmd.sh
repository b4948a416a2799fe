mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 170 -c 1 -f -o gpurun_out/prof_gemm_step \
    python bench.py --steps 1 --warmup 3 --step eager --no-cpu-baseline --no-e2e --no-all-configs > gpurun_out/ncu_gemm.log 2>&1
tail -2 gpurun_out/ncu_gemm.log
