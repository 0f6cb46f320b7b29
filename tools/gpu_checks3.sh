#!/bin/bash
mkdir -p gpurun_out
echo "== unet tests"; timeout 1500 python -m pytest tests/test_gpu_unet.py -m gpu -q -x -s > gpurun_out/test_unet.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/test_unet.log
bash tools/gpu_profile.sh
