mkdir -p gpurun_out
( time timeout 1500 python bench.py > gpurun_out/r40_default.json 2> gpurun_out/r40_default.err ) 2> gpurun_out/r40_time.txt
tail -2 gpurun_out/r40_default.err; cat gpurun_out/r40_time.txt
( time timeout 900 python bench.py --impl reference > gpurun_out/r40_reference.json 2> gpurun_out/r40_reference.err ) 2>> gpurun_out/r40_time.txt
tail -2 gpurun_out/r40_reference.err; tail -4 gpurun_out/r40_time.txt
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r40_default.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["vs_baseline"], d["clocks"], json.dumps(d["roofline"])[:300])
print(json.dumps(d["cpu_baseline"])[:400])
for k,v in d["other_models"].items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a in ("value","ms_per_step","error","e2e")})
r=json.loads(open("gpurun_out/r40_reference.json").read().strip().splitlines()[-1])
print({k:r[k] for k in ("impl","value","unit","ms_per_step")}, json.dumps(r.get("cpu_baseline"))[:300])
PY
