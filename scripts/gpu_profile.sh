#!/bin/bash
# ncu evidence for the bench step (run through gpurun on one B200). Outputs in gpurun_out/.
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline $@"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?" | tee -a gpurun_out/status.txt
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'hashgrid_fwd|hashgrid_bwd|adam_kernel' -s 6 -c 6 -f -o gpurun_out/prof_step $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?" | tee -a gpurun_out/status.txt
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'hashmlp_sweep|decoder2|gather_voxels|mse_kernel' -s 16 -c 9 -f -o gpurun_out/prof_rest $CMD > gpurun_out/ncu_full2.log 2>&1
echo "full capture 2 exit $?" | tee -a gpurun_out/status.txt
ls -la gpurun_out
