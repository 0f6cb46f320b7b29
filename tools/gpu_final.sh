#!/bin/bash
# round-end evidence on one GPU: full GPU test suite, smoke, both bench arms, ncu launch list of one window step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/test_gpu_all.log 2>&1; echo "pytest -m gpu rc=$?"; tail -3 gpurun_out/test_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_final.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_final.log | cut -c1-400
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "bench reference rc=$?"; tail -1 gpurun_out/bench_reference.log | cut -c1-300
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/profile_step.log 2>&1; echo "launch list rc=$?"
