mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --model autoint --big-tables 8x100000000 --no-cpu-baseline --no-other-models > gpurun_out/r34_autoint_big_n2.json 2> gpurun_out/r34.err
tail -4 gpurun_out/r34.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r34_autoint_big_n2.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"], d["windows_ms_per_step"], round(d["e2e"]["value"]), json.dumps(d.get("verified")), d["config"].get("parallelism") or d["run"].get("parallelism"))
print(json.dumps(d["op_ms"])[:900])
PY
