mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 > gpurun_out/final_n2.json 2> gpurun_out/final_n2.err
tail -2 gpurun_out/final_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/final_n2.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), json.dumps(d["verified"])[:200])
for k,v in d["other_models"].items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a in ("value","ms_per_step","error","scaling")})
PY
