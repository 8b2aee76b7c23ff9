set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "embed_bwd" 2>&1 | tail -5
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"route_|embed_segsum|embed_fixup" -s 27 -c 9 -f -o gpurun_out/r20_route python tools/embed_bench.py > gpurun_out/r20_ncu.log 2>&1
tail -3 gpurun_out/r20_ncu.log
ncu -i gpurun_out/r20_route.ncu-rep --page raw --csv > gpurun_out/r20_route_raw.csv 2>/dev/null
ls -la gpurun_out/
