#!/bin/bash
# The reference's three callers on the B200 backend (short runs). Outputs in gpurun_out/callers/.
mkdir -p gpurun_out/callers
timeout 600 python launcher.py --epochs 1 --batch_size 100000 --max_steps 60 --output_dir gpurun_out/callers > gpurun_out/callers/launcher.log 2>&1
echo "launcher exit $?"
tail -4 gpurun_out/callers/launcher.log
CK=$(ls gpurun_out/callers/lightning_logs/version_0/checkpoints/*.ckpt | head -1)
cat gpurun_out/callers/lightning_logs/version_0/scores.txt
timeout 300 python interp.py sweep --checkpoint "$CK" --model_class HashMLP --shape 352 352 6 57 --out gpurun_out/callers/interp57.nii.gz > gpurun_out/callers/interp.log 2>&1
echo "interp sweep exit $?"; tail -2 gpurun_out/callers/interp.log
timeout 300 python interp.py linear --out gpurun_out/callers/linear.nii.gz 2>&1 | tail -1
timeout 600 python test_script.py 2 > gpurun_out/callers/test_script.log 2>&1
echo "test_script exit $?"; tail -2 gpurun_out/callers/test_script.log
rm -f gpurun_out/callers/*.nii.gz gpurun_out/callers/lightning_logs/version_*/*.nii.gz gpurun_out/callers/lightning_logs/version_*/checkpoints/*.ckpt
