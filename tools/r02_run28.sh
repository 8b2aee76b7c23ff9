mkdir -p gpurun_out
nvidia-smi nvlink -gt d -i 0 | head -8
timeout 600 python -m pytest tests/test_parallel_gpu.py tests/test_peer_gpu.py -m gpu -q -x 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline --no-other-models > gpurun_out/r28_x_n2.json 2> gpurun_out/r28_n2.err
tail -3 gpurun_out/r28_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r28_x_n2.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["windows_ms_per_step"], json.dumps(d.get("nvlink")), json.dumps(d.get("verified")))
PY
