mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"attn_tc_bwd|attn_tc_fwd" -s 8 -c 2 -f -o gpurun_out/r48_attn python bench.py --model autoint --steps 1 --warmup 3 --windows 1 --no-cpu-baseline --no-graph --no-other-models > gpurun_out/r48_ncu.log 2>&1
tail -2 gpurun_out/r48_ncu.log
ncu -i gpurun_out/r48_attn.ncu-rep --page raw --csv > gpurun_out/r48_attn_raw.csv 2>/dev/null
ncu -i gpurun_out/r48_attn.ncu-rep --page source --csv > gpurun_out/r48_attn_src.csv 2>/dev/null
ls -la gpurun_out/r48*
