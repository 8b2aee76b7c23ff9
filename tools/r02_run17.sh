set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/peer_bench.py --out gpurun_out/r02_peer_bench_N2.json > gpurun_out/r02q_peer.log 2>&1
tail -30 gpurun_out/r02q_peer.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --no-other-models > gpurun_out/r02q_bench_n2.json 2> gpurun_out/r02q_bench_n2.err
timeout 600 python -m pytest tests/test_parallel_gpu.py -m gpu -q -x > gpurun_out/r02q_pytest_par.log 2>&1
tail -5 gpurun_out/r02q_pytest_par.log
