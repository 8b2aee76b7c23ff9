mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 > gpurun_out/r41_n8.json 2> gpurun_out/r41_n8.err ) 2> gpurun_out/r41_time.txt
tail -3 gpurun_out/r41_n8.err; cat gpurun_out/r41_time.txt
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r41_n8.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"], d["windows_ms_per_step"], round(d["e2e"]["value"]), json.dumps(d["verified"])[:400])
for k,v in d["other_models"].items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a in ("value","ms_per_step","error","scaling")})
print({k:round(v["ms"],4) for k,v in d["op_ms"].items()})
PY
