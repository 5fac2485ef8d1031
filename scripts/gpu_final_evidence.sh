#!/bin/bash
# Round-end evidence on one B200: GPU tests, the default bench, ncu launch list and ncu --set full captures of the hot kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc $?"
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc $?"
STEP="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-workloads --no-traffic-probe"
timeout 300 $STEP --no-infer > gpurun_out/plain_step.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 300 --csv --log-file gpurun_out/launches_step.csv $STEP --no-infer > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'hashdecoder_mma' -s 8 -c 2 -f -o gpurun_out/prof_fused $STEP --no-infer > gpurun_out/ncu_fused.log 2>&1; echo "ncu fused rc $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'SweepCoordsAxis0|adam_kernel' -s 40 -c 2 -f -o gpurun_out/prof_sweep $STEP > gpurun_out/ncu_sweep.log 2>&1; echo "ncu sweep rc $?"
SIREN="python bench.py --workload siren_wide --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-infer --no-traffic-probe"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'siren_tc_layer_kernel|siren_tc_wgrad_kernel' -s 60 -c 4 -f -o gpurun_out/prof_siren $SIREN > gpurun_out/ncu_siren.log 2>&1; echo "ncu siren rc $?"
ls -la gpurun_out/*.ncu-rep
