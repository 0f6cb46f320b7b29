#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "attention" > gpurun_out/test_ops_attention.log 2>&1; echo "attention rc=$?"; tail -3 gpurun_out/test_ops_attention.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['by_kind_ms'], d['roofline']['by_kind_tflops'])"
timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/profile_step.log 2>&1; echo "launch list rc=$?"
timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_fwd -s 4 -c 1 -o gpurun_out/prof_attn -f python tools/profile_step.py > gpurun_out/prof_attn.log 2>&1; echo "attn capture rc=$?"
timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_umma -s 28 -c 8 -o gpurun_out/prof_gemm -f python tools/profile_step.py > gpurun_out/prof_gemm.log 2>&1; echo "gemm capture rc=$?"
timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gn_ -s 6 -c 4 -o gpurun_out/prof_gn -f python tools/profile_step.py > gpurun_out/prof_gn.log 2>&1; echo "gn capture rc=$?"
