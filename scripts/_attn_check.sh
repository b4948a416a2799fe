timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_step.py tests/test_gpu_train_loop.py tests/test_gpu_layers.py -q -x 2>&1 | tail -3
timeout 300 python bench.py --no-all-configs --no-cpu-baseline --steps 10 --warmup 3 2>gpurun_out/attn_check.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],2), d.get('e2e',{}).get('value'), d.get('gpu_launches'), d.get('our_kernel_ms_per_step'), d['clocks'])"
tail -2 gpurun_out/attn_check.err
