#!/bin/bash
# One GPU.  ncu captures behind profiles/r02_*.md (numbers taken under ncu are never bench values).
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
# launch lists (eager, every kernel its own launch; --windows 1 keeps the run short)
for M in xdeepfm deepfm; do
  timeout 300 $NCU --metrics gpu__time_duration.sum -c 2200 --csv --log-file $O/r02_${M}_launches.csv \
      python bench.py --model $M --steps 2 --warmup 3 --windows 1 --no-cpu-baseline --no-graph --no-other-models > $O/ncu_${M}.log 2>&1
done
# full captures: the CIN kernels of one xDeepFM step (skip the 3 warm-up + profiled-eager steps' first instances)
timeout 400 $NCU --set full --import-source on -k regex:"cin_(fwd|da|dw2?)_tc|cin_last" -s 24 -c 9 -f -o $O/r02_cin \
    python bench.py --steps 1 --warmup 3 --windows 1 --no-cpu-baseline --no-graph --no-other-models > $O/ncu_cin.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:"embed_fwd_vec|embed_reduce|embed_keys|embed_adam|pad1|Onesweep" -s 30 -c 12 -f -o $O/r02_embed \
    python bench.py --model deepfm --steps 1 --warmup 3 --windows 1 --no-cpu-baseline --no-graph --no-other-models > $O/ncu_embed.log 2>&1
ls -la $O/*.ncu-rep
for R in r02_cin r02_embed; do
  python tools/ncu_summary.py $O/$R.ncu-rep > $O/${R}_summary.md 2>/dev/null
  ncu -i $O/$R.ncu-rep --page raw --csv > $O/${R}_raw.csv 2>/dev/null
done
