mkdir -p gpurun_out
ncu --query-metrics 2>/dev/null | grep -i "nvl" | head -20 > gpurun_out/r29_nvl_metrics.txt; cat gpurun_out/r29_nvl_metrics.txt | cut -c1-120
KON_PEER_TIMEOUT_MS=120000 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 --no-python tools/r02_peer_ncu_rank.sh --gpus 2 --steps 2 --warmup 3 --windows 1 --no-cpu-baseline --no-other-models --no-graph > gpurun_out/r29_out.log 2> gpurun_out/r29_err.log
tail -5 gpurun_out/r29_err.log
grep -v "^==" gpurun_out/r02_peer_nvl.csv | cut -d, -f5,13-16 | tail -40
