#!/bin/bash
# N-GPU checks: frame-sharded parity test, then the N-rank bench (replicas line + `sharded` object)
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x > gpurun_out/test_sharded.log 2>&1; echo "sharded test rc=$?"; tail -3 gpurun_out/test_sharded.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_n$N.log | cut -c1-600
