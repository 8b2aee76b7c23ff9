mkdir -p gpurun_out
for M in deepfm dcn autoint; do
timeout 300 ncu --clock-control none --metrics gpu__time_duration.sum -c 2200 --csv --log-file gpurun_out/r35_${M}_launches.csv python bench.py --model $M --steps 2 --warmup 3 --windows 1 --no-cpu-baseline --no-graph --no-other-models > gpurun_out/r35_ncu_$M.log 2>&1
python tools/launch_list.py gpurun_out/r35_${M}_launches.csv gpurun_out/r35_${M}_step.csv > /dev/null 2>&1
done
ls -la gpurun_out/r35*
