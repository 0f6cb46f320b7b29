#!/bin/bash
# 2-GPU checks: frame-sharded parity test, then replicas and sharded bench at N=2
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x -s > gpurun_out/test_sharded.log 2>&1; echo "sharded test rc=$?"; tail -15 gpurun_out/test_sharded.log
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_rep$N.log 2>&1; echo "replicas rc=$?"; tail -1 gpurun_out/bench_rep$N.log | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_shard$N.log 2>&1; echo "sharded rc=$?"; tail -3 gpurun_out/bench_shard$N.log | cut -c1-400
