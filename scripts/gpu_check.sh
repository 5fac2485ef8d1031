#!/bin/bash
# Run on the B200 box through gpurun: GPU tests, smoke, bench; results land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/status.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/status.txt
timeout 900 python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" | tee -a gpurun_out/status.txt
tail -5 gpurun_out/pytest_gpu.log
tail -3 gpurun_out/smoke.log
cat gpurun_out/bench.json
tail -5 gpurun_out/bench.err
