# 2-GPU check of the final code: DP gradient test, strong-scaling bench line (256 graphs over 2 GPUs)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dp.py -x -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/bench_2gpu.err | tail -1 > gpurun_out/bench_2gpu.json
tail -2 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_2gpu.json"))
print(round(d["value"], 1), "graphs/s", round(d["ms_per_step"], 2), "ms/step  e2e", d.get("e2e") and round(d["e2e"]["value"], 1), "dp_grad_rel_err", d.get("dp_grad_rel_err"), d["config"]["step_execution"], d["clocks"])
PY
