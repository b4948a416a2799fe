# 8-GPU strong-scaling line of the final code (one batch of 256 graphs over 8 ranks) + the 1-GPU line on the same box
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 2>gpurun_out/bench_8gpu.err | tail -1 > gpurun_out/bench_8gpu.json
tail -2 gpurun_out/bench_8gpu.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-all-configs --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_1gpu_same_box.json
python - <<'PY'
import json
a = json.load(open("gpurun_out/bench_8gpu.json")); b = json.load(open("gpurun_out/bench_1gpu_same_box.json"))
print("8 GPUs:", round(a["value"], 1), "graphs/s", round(a["ms_per_step"], 2), "ms/step  e2e", a.get("e2e") and round(a["e2e"]["value"], 1), "dp_grad_rel_err", a.get("dp_grad_rel_err"), a["config"]["step_execution"], a["clocks"])
print("1 GPU :", round(b["value"], 1), "graphs/s", round(b["ms_per_step"], 2), "ms/step  -> speed-up", round(a["value"] / b["value"], 2), "e2e", round(a["e2e"]["value"] / b["e2e"]["value"], 2))
PY
