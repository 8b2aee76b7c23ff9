mkdir -p gpurun_out
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"embed_segsum|embed_fixup|embed_adam_kernel|embed_fwd_vec" -s 6 -c 5 -f -o gpurun_out/r55_embed_step python bench.py --model deepfm --steps 1 --warmup 3 --windows 1 --no-cpu-baseline --no-graph --no-other-models > gpurun_out/r55_ncu.log 2>&1
tail -2 gpurun_out/r55_ncu.log
ncu -i gpurun_out/r55_embed_step.ncu-rep --page raw --csv > gpurun_out/r55_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r55_raw.csv')))
h=rows[0]
want=["Kernel Name","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","dram__throughput.avg.pct_of_peak_sustained_elapsed","sm__warps_active.avg.pct_of_peak_sustained_active","smsp__issue_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","smsp__inst_executed.sum"]
idx=[h.index(w) for w in want if w in h]
for r in rows[2:]: print(" | ".join(r[i][:44] for i in idx))
PY
