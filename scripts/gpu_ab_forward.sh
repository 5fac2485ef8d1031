#!/bin/bash
# A/B of the encoder+decoder forward fusion and the tensor-core sweep kernel (one B200).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x -k "hashdecoder or sweep or smoke or hashmlp" > gpurun_out/ab_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/ab_pytest.log
for cfg in "1 0" "0 1"; do
  set -- $cfg
  if [ "$2" = "1" ]; then export MRI_SWEEP_CUDA_CORES=1; else unset MRI_SWEEP_CUDA_CORES; fi
  MRI_FUSED_FORWARD=$1 MRI_BENCH_NO_CLOCKS=1 timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/ab_fwd$1_cc$2.json 2> gpurun_out/ab_fwd$1_cc$2.err
  echo "fused_forward=$1 sweep_cuda_cores=$2 exit $?"
  python - "$1" "$2" <<'PY'
import json, sys
for l in open(f"gpurun_out/ab_fwd{sys.argv[1]}_cc{sys.argv[2]}.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("  ms/step", round(d["ms_per_step"], 4), "Mcoords/s", round(d["value"] / 1e6, 1), "launches", d["gpu_launches"],
              "infer Gvox/s", round(d["infer"]["value"] / 1e9, 3), d["infer"]["ms_per_sweep"],
              {k: round(v["ms"], 4) for k, v in d["kernels"].items()})
PY
done
