"""Build libmri_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m mri_interpolation_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels with the repo
snapshot to the GPU box.  No torch headers are involved: the ABI is plain C (include/mri_b200.h).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libmri_b200.so")
SOURCES = ["core.cu", "hashgrid.cu", "dense.cu", "optim.cu", "sweep.cu", "siren_tc.cu", "decoder.cu", "decoder_k16.cu", "decoder_k32.cu", "decoder_k64.cu", "siren_edge.cu", "hashdecoder_fwd.cu",
           "hashdecoder_bwd.cu", "metrics.cu", "probe.cu", "hashdecoder_fwd_geo.cu", "hashdecoder_bwd_geo.cu", "hashdecoder_step.cu"]
HEADERS = ["common.cuh", "hash_device.cuh", "mma_device.cuh", "grid_device.cuh", "hashdecoder.cuh", "hashdecoder_fwd_impl.cuh",
           "hashdecoder_bwd_impl.cuh", "decoder_impl.cuh", "../../include/mri_b200.h"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.isfile(os.path.join(CSRC, s))]


def up_to_date() -> bool:
    if not os.path.isfile(LIB_PATH):
        return False
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    return all(os.path.getmtime(d) <= t for d in deps if os.path.isfile(d))


def _headers_mtime() -> float:
    deps = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    return max(os.path.getmtime(d) for d in deps if os.path.isfile(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every source whose object is older than the source or any header (all in parallel), then link."""
    if not force and up_to_date():
        return LIB_PATH
    nvcc = _nvcc()
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    hdr_t = _headers_mtime()
    for src in sources():
        obj = os.path.join(HERE, "build", os.path.basename(src) + ".o")
        objs.append(obj)
        if not force and os.path.isfile(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), hdr_t):
            continue
        log_path = obj + ".log"
        cmd = [nvcc, *NVCC_FLAGS, "-Xptxas", "-v", "-c", src, "-o", obj]
        procs.append((src, log_path, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = []
    for src, log_path, p in procs:
        out, _ = p.communicate()
        with open(log_path, "w") as f:
            f.write(out)
        if p.returncode != 0:
            failed.append((src, out))
            obj = log_path[:-4]
            if os.path.isfile(obj):
                os.remove(obj)
        elif verbose:
            sys.stderr.write(f"==== {os.path.basename(src)}\n{out}\n")
    # one combined register/spill log, in source order (per-object logs persist across incremental builds)
    with open(os.path.join(HERE, "build", "ptxas.log"), "w") as f:
        for obj in objs:
            if os.path.isfile(obj + ".log"):
                f.write(f"==== {os.path.basename(obj)[:-2]}\n{open(obj + '.log').read()}\n")
    if failed:
        for src, out in failed:
            sys.stderr.write(f"==== {os.path.basename(src)}\n{out}\n")
        raise RuntimeError("nvcc failed, see log above")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-o", LIB_PATH,
            *objs]
    subprocess.run(link, check=True)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
