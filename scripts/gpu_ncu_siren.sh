#!/bin/bash
# ncu evidence for the tensor-core SIREN kernels (tensor pipe utilisation). Outputs in gpurun_out/.
CMD="python scripts/microbench.py sirenlayer"
$CMD > gpurun_out/plain_sirenlayer.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'siren_tc_layer|siren_tc_wgrad' -s 6 -c 4 -f -o gpurun_out/prof_siren $CMD > gpurun_out/ncu_siren_full.log 2>&1
echo "siren capture exit $?"
