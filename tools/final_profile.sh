#!/bin/bash
# One GPU.  ncu captures behind the summaries under profiles/ (numbers under ncu are never bench values).
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
# launch list of the xDeepFM step (eager, so every kernel is its own launch)
timeout 240 $NCU --metrics gpu__time_duration.sum -c 1400 --csv --log-file $O/r01b_xdeepfm_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > $O/ncu_x.log 2>&1
# full captures of the kernels this round touched
timeout 200 $NCU --set full --import-source on -k regex:"cin_(fwd|da|dw2?)_tc|cin_last" -s 9 -c 9 -f -o $O/r01b_cin \
    python tools/cin_bench.py 65536 bwd > $O/ncu_cin.log 2>&1
timeout 200 $NCU --set full --import-source on -k regex:attn_tc -s 4 -c 2 -f -o $O/r01b_attn \
    python tools/attn_bench.py 65536 bf16 > $O/ncu_attn.log 2>&1
timeout 240 $NCU --set full --import-source on -k regex:"cross_|head_" -s 16 -c 8 -f -o $O/r01b_cross \
    python bench.py --model dcn --steps 1 --warmup 3 --no-cpu-baseline --no-graph > $O/ncu_cross.log 2>&1
timeout 240 $NCU --set full --import-source on -k regex:"embed_fwd_vec|embed_reduce|embed_adam" -s 12 -c 6 -f -o $O/r01b_embed \
    python bench.py --model deepfm --steps 1 --warmup 3 --no-cpu-baseline --no-graph > $O/ncu_embed.log 2>&1
ls -la $O/*.ncu-rep
