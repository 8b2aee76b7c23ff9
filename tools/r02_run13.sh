set -x
nvidia-smi nvlink -gt d -i 0 | head -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/peer_bench.py --out gpurun_out/r02_peer_bench_N2.json 2>&1 | tail -5
timeout 600 python -m pytest tests/test_models_gpu.py -m gpu -q 2>&1 | tail -5 > gpurun_out/r02m_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 > gpurun_out/r02m_bench_n2.json 2> gpurun_out/r02m_bench_n2.err
tail -c 600 gpurun_out/r02m_bench_n2.json
