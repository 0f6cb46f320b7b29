#!/bin/bash
mkdir -p gpurun_out
echo "== unet tests"; timeout 1200 python -m pytest tests/test_gpu_unet.py -m gpu -q -x -s > gpurun_out/test_unet.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/test_unet.log
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/smoke.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/bench.log
