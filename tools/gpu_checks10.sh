#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x > gpurun_out/test_ops.log 2>&1; echo "ops rc=$?"; tail -3 gpurun_out/test_ops.log
for a in 0 32; do D4D_GEMM_ABLATE=$a timeout 120 python tools/ablate_gemm.py --small; done 2>&1 | sed 's/cyc\/kb//g' | tee gpurun_out/ablate_gemm_small2.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['by_kind_ms'], d['roofline']['by_kind_tflops'])"
timeout 900 python -m pytest tests/test_gpu_unet.py -m gpu -q -x > gpurun_out/test_unet.log 2>&1; echo "unet rc=$?"; tail -3 gpurun_out/test_unet.log
