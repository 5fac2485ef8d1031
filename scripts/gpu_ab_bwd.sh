#!/bin/bash
# A/B of the fused backward kernel's merge depth (MRI_BWD_MERGE_LEVELS) on one B200; prints ms/step and kernel ms.
mkdir -p gpurun_out
for m in 8 4 0; do
  MRI_BWD_MERGE_LEVELS=$m timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-infer --no-e2e --no-workloads > gpurun_out/ab_merge$m.json 2> gpurun_out/ab_merge$m.err
  python - <<PY
import json
d = json.load(open("gpurun_out/ab_merge$m.json"))
k = d["kernels"]
print("merge $m: step %.4f ms  bwd %.4f (frac %.3f)  fwd %.4f (frac %.3f)  adam %.4f" % (d["ms_per_step"], k["hashdecoder_bwd"]["ms"], k["hashdecoder_bwd"]["frac"], k["hashdecoder_fwd"]["ms"], k["hashdecoder_fwd"]["frac"], k["adam_step"]["ms"]))
PY
done
