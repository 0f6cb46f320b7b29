#!/bin/bash
mkdir -p gpurun_out
for grp in "gemm or conv3x3" attention "groupnorm or layernorm"; do
  timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "$grp" > "gpurun_out/test_ops_${grp// /_}.log" 2>&1
  echo "$grp rc=$?"; tail -4 "gpurun_out/test_ops_${grp// /_}.log"
done
echo "== unet tests"; timeout 1500 python -m pytest tests/test_gpu_unet.py -m gpu -q -x -s > gpurun_out/test_unet.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/test_unet.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/bench.log
