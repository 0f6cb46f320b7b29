#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/microbench.py --pipe-only 2>&1 | tee gpurun_out/microbench_pipe.txt
for a in 0 16 3 19; do D4D_GEMM_ABLATE=$a timeout 120 python tools/ablate_gemm.py; done 2>&1 | tee gpurun_out/ablate_gemm2.txt
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "gemm or conv3x3" 2>&1 | tail -2
