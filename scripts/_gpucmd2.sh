# 2-GPU check: DP test, strong-scaling bench lines (graph with captured all-reduce; fallback mode), weak line
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dp.py -x -q 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/bench_2gpu.err | tail -1 > gpurun_out/bench_2gpu.json
tail -3 gpurun_out/bench_2gpu.err
DCB200_GRAPH_ALLREDUCE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e 2>gpurun_out/bench_2gpu_onegraph.err | tail -1 > gpurun_out/bench_2gpu_onegraph.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --step eager 2>gpurun_out/bench_2gpu_eager.err | tail -1 > gpurun_out/bench_2gpu_eager.json
python - <<'PY'
import json
for f in ("bench_2gpu", "bench_2gpu_onegraph", "bench_2gpu_eager"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"], 1), round(d["ms_per_step"], 2), d.get("e2e") and round(d["e2e"]["value"], 1), d.get("dp_grad_rel_err"), d["config"]["step_execution"], d["gpu_launches_per_step"])
    except Exception as e:
        print(f, "FAILED", e)
PY
for mode in graph eager; do timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --global-batch 64 --steps 20 --warmup 3 --no-e2e --step $mode 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('64 graphs on 2 GPUs (32 per rank, as at 8 GPUs), step=$mode:', round(d['ms_per_step'],2), 'ms/step', round(d['value'],1), 'graphs/s', d['config']['step_execution'])"; done
