set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x -k "embed" 2>&1 | tail -5
timeout 300 python tools/embed_bench.py > gpurun_out/r31_embed_uniform.json 2> gpurun_out/r31_embed.err
timeout 300 python tools/embed_bench.py zipf > gpurun_out/r31_embed_zipf.json 2>> gpurun_out/r31_embed.err
cat gpurun_out/r31_embed_uniform.json gpurun_out/r31_embed_zipf.json; tail -5 gpurun_out/r31_embed.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r31_embed_launches.csv python tools/embed_bench.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[l for l in open('gpurun_out/r31_embed_launches.csv') if not l.startswith('==')]
r=list(csv.reader(rows)); h=r[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
for x in r[-11:]:
    print(x[ki][:60], x[vi])
PY
