python scripts/acc_lab.py 2>&1 | tail -12
DCB200_GEMM=fp32 python scripts/acc_lab.py 2>&1 | tail -4
DCB200_ATTENTION=torch python scripts/acc_lab.py 2>&1 | tail -4
DCB200_DECODER=torch python scripts/acc_lab.py 2>&1 | tail -4
DCB200_LOSS=torch python scripts/acc_lab.py 2>&1 | tail -4
DCB200_ATTENTION=torch DCB200_DECODER=torch DCB200_LOSS=torch python scripts/acc_lab.py 2>&1 | tail -4
DCB200_ATTENTION=torch DCB200_DECODER=torch DCB200_LOSS=torch DCB200_GEMM=fp32 python scripts/acc_lab.py 2>&1 | tail -4
timeout 300 python -m pytest tests/test_gpu_step.py tests/test_gpu_train_loop.py -x -q -o faulthandler_timeout=120 2>&1 | tail -30
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "chain or relu_bwd" 2>&1 | tail -8
timeout 300 python scripts/k1_chain_lab.py --out gpurun_out/k1_chain_lab.json 2>&1 | tail -14
