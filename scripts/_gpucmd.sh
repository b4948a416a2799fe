echo "== tests with DCB200_BRANCH_STREAMS=1"
DCB200_BRANCH_STREAMS=1 timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_train_loop.py tests/test_gpu_layers.py -q -x 2>&1 | tail -2
for bs in 1 0 1 0; do for b in 32 256; do DCB200_BRANCH_STREAMS=$bs python bench.py --global-batch $b --steps 20 --no-all-configs --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('branch_streams=$bs graphs $b:', round(d['ms_per_step'],3), 'ms/step')"; done; done
