mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"route_scatter|route_heads|embed_segsum" -s 9 -c 4 -f -o gpurun_out/r24_route python tools/embed_bench.py > gpurun_out/r24_ncu.log 2>&1
tail -3 gpurun_out/r24_ncu.log
ncu -i gpurun_out/r24_route.ncu-rep --page raw --csv > gpurun_out/r24_route_raw.csv 2>/dev/null
ncu -i gpurun_out/r24_route.ncu-rep --page source --csv > gpurun_out/r24_route_src.csv 2>/dev/null
ls -la gpurun_out/ | tail -5
