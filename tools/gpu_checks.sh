#!/bin/bash
# Run on the GPU box (via gpurun).  Each group runs in its own process under `timeout` so that a trapped
# kernel (poisoned context) or a hang cannot take the other groups down.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== probe" ; timeout 300 python tools/probe_umma.py > gpurun_out/probe.txt 2>&1; echo "probe rc=$?"; tail -12 gpurun_out/probe.txt
for grp in umma_descriptor gemm conv3x3 attention groupnorm layernorm; do
  echo "== $grp"
  timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "$grp" > gpurun_out/test_$grp.log 2>&1
  echo "$grp rc=$?"; tail -15 gpurun_out/test_$grp.log
done
