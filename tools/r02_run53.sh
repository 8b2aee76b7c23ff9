mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"cin_last_dw_kernel|cin_last_da_tc|cin_last_pool" -s 3 -c 3 -f -o gpurun_out/r53_last python bench.py --steps 1 --warmup 3 --windows 1 --no-cpu-baseline --no-graph --no-other-models > gpurun_out/r53_ncu.log 2>&1
tail -2 gpurun_out/r53_ncu.log
ncu -i gpurun_out/r53_last.ncu-rep --page raw --csv > gpurun_out/r53_raw.csv 2>/dev/null
ncu -i gpurun_out/r53_last.ncu-rep --page source --csv > gpurun_out/r53_src.csv 2>/dev/null
ls -la gpurun_out/r53*
