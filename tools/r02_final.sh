# end-of-round check on one GPU: smoke(), the GPU test suite, the default bench line
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 1500 python bench.py > gpurun_out/final_default.json 2> gpurun_out/final_default.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/final_default.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["clocks"], d["roofline"]["frac"], d["gpu_launches"])
for k,v in d["other_models"].items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a in ("value","ms_per_step","error")})
print({k:(round(v["frac"],3), round(v["ms"],4)) for k,v in d["op_stats"].items()})
PY
