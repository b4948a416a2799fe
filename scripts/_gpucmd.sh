timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.json; tail -5 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_default.json"))
drop=("note","kernel","workload","attention","e2e_inputs","sample","traffic_source","kind")
print({k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk not in drop}) for k,v in d.items() if k!="config"})
print(d["config"]["step_execution"])
PY
