#!/bin/bash
# launch list of one window step + full ncu captures of the top kernels (1 GPU; numbers under ncu are never bench values)
mkdir -p gpurun_out
timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/profile_step.log 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/profile_step.log; wc -l gpurun_out/launches.csv
timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_fwd -c 3 \
    -o gpurun_out/prof_attn -f python tools/profile_step.py > gpurun_out/prof_attn.log 2>&1
echo "attn capture rc=$?"
timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_umma -s 30 -c 4 \
    -o gpurun_out/prof_gemm -f python tools/profile_step.py > gpurun_out/prof_gemm.log 2>&1
echo "gemm capture rc=$?"
ls -la gpurun_out/*.ncu-rep
