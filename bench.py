#!/usr/bin/env python
"""Benchmark of the coordinate-network hot path (BASELINE.json metric: training coords/s, plus inference
voxels/s, with % of roofline and the reference's CPU path timed beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch-log2 B] [--workload NAME]

Headline workload (configs[1] of BASELINE.json, named in `config.workload`): hash-grid encoder G4
(config/hash_config.json: L=16, F=2, T=2^19, base 16, x1.4 -> finest 2489; 15 279 648 table params)
+ 2-layer GELU decoder (64 hidden), fitted to the sample ankle volume (352x352x6x15, x,y,z,t coords),
fp32, Adam lr 5e-3, 2^B coordinates per step PER GPU (weak scaling; N>1 exchanges the flat gradient arena once
per step inside the sharded Adam kernel).  One step = 5 kernels: synthesise coords / gather intensities from the
step's voxel indices -> hash encode + decoder (one kernel) -> MSE -> decoder backward + hash scatter (one kernel)
-> fused Adam [with the reduce-scatter / all-gather over NVLink at N>1].  The voxel indices are shuffled epochs
(a permutation of the volume cut into batches, as the reference's DataLoader does), each batch arranged in locality
order (axis-0 index fastest: same sets, see datamodules.ShuffledEpochs); they are drawn before the timed region and
their amortised cost is reported as `sampler_ms_per_batch`.

value : device-resident inputs, CUDA-event timing of EXACTLY --steps steps, max over ranks.  `sustained` repeats the
        same steps for >= 0.6 s (the clock samples cover the timed region plus that window).
e2e   : same step through the public LightningModule API with HOST (pinned) batches: H2D copy of the
        batch and D2H read of the loss inside the timed region.
roofline : dominant kernel (decoder backward fused with the hash-grid scatter) timed alone with CUDA events, L2
        flushed between launches; algorithmic bytes / duration against MEASURED_PEAKS.json's HBM copy bandwidth.
workloads : short legs of the other BASELINE configs (1: SIREN on the ankle volume, 4: synthetic 256^3x32 hash,
        5: wide SIREN 8x1024 on the tcgen05 path), each with its own roofline and clocks.
cpu_baseline / --impl reference : the reference's own classes (baseline/_ref, unmodified) on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

G4 = dict(n_levels=16, n_features_per_level=2, log2_hashmap_size=19, base_resolution=16, finest_resolution=2489)
SAMPLE = os.path.join(ROOT, "data", "sample_ankle_dyn_mri.nii.gz")
SWEEP_SHAPE = (352, 352, 6, 29)  # 2x time up-sampling of the sample volume (config 3)
PRIMING_STEPS = 15  # allocator priming before the W warm-up steps (reported in config)
HASH_BYTES_PER_COORD = 4 * 4 + 16 * 16 * 2 * 4 + 16 * 2 * 4  # 4D + L*2^D*F*4 + L*F*4 = 2192
# Fallback for roofline.traffic when the live ncu probe (traffic_probe below) cannot run: dram__bytes_read.sum +
# dram__bytes_write.sum and L2 reduction / L1 load sectors per launch at 2^19 locality-ordered coords, from the
# ncu --set full capture committed as profiles/r02_ncu_full_fused_hashdecoder.csv
NCU_DRAM_BYTES = {"hashdecoder_bwd": 142.2e6 + 149.4e6, "hashdecoder_fwd": 87.5e6 + 56.7e6}
NCU_RED_SECTORS = {"hashdecoder_bwd": 74.97e6}
NCU_SOURCE = "profiles/r02_ncu_full_fused_hashdecoder.csv"
PROBE_METRICS = ("dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors_srcunit_tex_op_red.sum",
                 "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum")
SUSTAIN_SECONDS = 0.6


WORKLOADS = {
    # name: (description, default log2 batch per GPU)
    "ankle_hash": ("hash-grid G4 (L16 F2 T2^19 base16 finest2489, 15.28M table params) + 2x64 GELU decoder fitted to "
                   "sample_ankle_dyn_mri.nii.gz (352x352x6x15, xyzt coords), Adam lr 5e-3", 19),
    "synthetic_hash": ("config 4: hash-grid G4 + 2x64 GELU decoder on a synthetic 256^3 x 32-frame volume (536.9M voxels, "
                       "sum of separable sinusoids + noise, generated on device), i.i.d. voxel indices per step, Adam lr 5e-3", 19),
    "siren_wide": ("config 5: SirenNet 3 -> 1024 x 8 -> 1 (w0 30) on a synthetic 512^3 volume, coords in [-1,1], "
                   "tensor-core split-precision mode bf16x3 (fp32 parity), Adam lr 1e-4", 17),
    "siren_ankle": ("config 1: SirenNet 4 -> 256 x 5 -> 1 (w0 30) on sample_ankle_dyn_mri.nii.gz, tensor-core "
                    "split-precision mode bf16x3 (fp32 parity), Adam lr 1e-4", 18),
}
EXTRA_LEGS = ("siren_wide", "synthetic_hash", "siren_ankle")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch-log2", type=int, default=None, help="log2 coordinates per step per GPU (default per workload)")
    ap.add_argument("--workload", default="ankle_hash", choices=list(WORKLOADS),
                    help="ankle_hash = BASELINE configs[1] (default, the driver's bench line); the others are configs 1/4/5")
    ap.add_argument("--cpu-batch-log2", type=int, default=19, help="coords per CPU step of the reference arm / cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-infer", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-workloads", action="store_true", help="skip the short legs of configs 1/4/5 after the headline")
    ap.add_argument("--no-traffic-probe", action="store_true",
                    help="skip the nested ncu run that measures roofline.traffic (DRAM bytes, L2 reduction sectors) on this box")
    ap.add_argument("--probe-steps", type=int, default=0, help=argparse.SUPPRESS)  # child of the traffic probe: N plain steps, no output
    args = ap.parse_args()
    if args.batch_log2 is None:
        args.batch_log2 = WORKLOADS[args.workload][1]
    return args


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return p, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (subprocess writing to a temp file:
    no Python reader thread competing with the launch loop for the GIL)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path, self.skip = index, None, None, 0

    def start(self):
        if os.environ.get("MRI_BENCH_NO_CLOCKS") == "1":
            return
        try:
            import tempfile
            fd, self.path = tempfile.mkstemp(prefix="mri_clocks_", suffix=".csv")
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", os.environ.get("MRI_BENCH_CLOCK_MS", "20")],
                                         stdout=fd, stderr=subprocess.DEVNULL)
            os.close(fd)
        except Exception:  # noqa: BLE001
            self.proc = None

    def mark(self):
        """Start of the window of interest: samples written before this call are dropped by stop()."""
        self.skip = 0
        if self.proc is not None:
            try:
                self.skip = len(open(self.path).read().splitlines())
            except Exception:  # noqa: BLE001
                self.skip = 0

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        try:
            lines = open(self.path).read().splitlines()[self.skip:]
            os.unlink(self.path)
        except Exception:  # noqa: BLE001
            lines = []
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------- reference arm
def reference_step_fn(batch_log2: int):
    """One training step of the reference's OWN classes on the CPU: `encoding.MultiResHashGrid` (unmodified, imported from
    baseline/_ref) in the G4 geometry + the notebook's Linear -> GELU decoder blocks (nb cell 37; the shipped
    HashMLP.forward calls a ModuleList and cannot run) + F.mse_loss + torch.optim.Adam(lr 5e-3), batches cut from the
    sample volume exactly like MriImage does (linspace coords, min-max intensities).  Returns (step, n, kind)."""
    import torch.nn.functional as F
    from mri_interpolation_b200 import nifti
    vol = torch.from_numpy(nifti.load(SAMPLE).get_fdata(np.float32))
    pix = vol.flatten()
    pix = ((pix - pix.min()) / (pix.max() - pix.min())).unsqueeze(-1)
    axes = [torch.linspace(0, 1, s) for s in vol.shape]
    shape = torch.tensor(vol.shape)
    n = 1 << batch_log2
    gen = torch.Generator().manual_seed(1337)
    torch.manual_seed(1337)
    from baseline import load_reference
    if load_reference.available():
        ref_encoding, _ = load_reference.load()
        encoder = ref_encoding.MultiResHashGrid(4, **G4)
        decoder = torch.nn.Sequential(torch.nn.Linear(32, 64), torch.nn.GELU(), torch.nn.Linear(64, 1), torch.nn.GELU())
        params = list(encoder.parameters()) + list(decoder.parameters())
        forward = lambda x: decoder(encoder(x))  # noqa: E731
        kind = "reference"
    else:  # no copy of the reference on this box: the oracle port of the same arithmetic
        from oracle import networks
        p, levels = networks.hashmlp_init(dim_in=4, dim_hidden=64, dim_out=1, n_layers=2, **G4)
        p = {k: v.requires_grad_() for k, v in p.items() if not k.startswith("layers.")}
        params = list(p.values())
        forward = lambda x: networks.hashmlp_forward(x, p, levels, 2, False)  # noqa: E731
        kind = "port"
    opt = torch.optim.Adam(params, lr=5e-3)

    def step():
        idx = torch.randint(0, pix.shape[0], (n,), generator=gen)
        rem, cols = idx.clone(), []
        for d in range(3, -1, -1):
            cols.append(axes[d][rem % shape[d]])
            rem = rem // shape[d]
        x = torch.stack(cols[::-1], dim=-1)
        opt.zero_grad()
        loss = F.mse_loss(pix[idx], forward(x))
        loss.backward()
        opt.step()
        return float(loss.detach())

    return step, n, kind


def time_reference(batch_log2: int, steps: int, warmup: int):
    torch.set_num_threads(os.cpu_count() or 1)
    step, n, kind = reference_step_fn(batch_log2)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return n / dt, dt, n, kind


def cpu_sample_text(kind, steps, log2, dt):
    what = ("the reference's own encoding.MultiResHashGrid (baseline/_ref, unmodified) + nb-cell-37 Linear/GELU decoder + "
            "F.mse_loss + torch.optim.Adam" if kind == "reference" else
            "oracle port of the reference's PyTorch CPU path: hash G4 + decoder + MSE + torch.optim.Adam")
    return f"{steps} steps of 2^{log2} coords of the same training step ({what}), {dt * 1e3:.0f} ms/step"


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    if args.workload != "ankle_hash":
        print(json.dumps({"impl": "reference", "unavailable": "the CPU reference arm is implemented for the default "
                          "workload (ankle_hash, BASELINE configs[1]) only"}))
        return
    steps, warm = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    value, dt, n, kind = time_reference(args.cpu_batch_log2, steps, warm)
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": "train_coords_per_s", "value": value, "unit": "coords/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "sample_ankle_dyn_mri.nii.gz (bundled) + random-init weights",
        "config": workload_config(args, "ankle_hash", args.cpu_batch_log2, 1, cpu_sample=n),
        "cpu_baseline": {"value": value, "unit": "coords/s", "cores": cores, "kind": kind,
                         "sample": cpu_sample_text(kind, steps, args.cpu_batch_log2, dt)},
        "e2e": {"value": value, "unit": "coords/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, name, batch_log2, world, cpu_sample=None):
    cfg = {"workload": WORKLOADS[name][0], "workload_name": name,
           "batch_per_gpu": 1 << batch_log2, "global_batch": (1 << batch_log2) * world,
           "parallelism": f"dp{world}" + (" (fused reduce-scatter + Adam + all-gather kernel over NVLink peer memory; "
                                          "NCCL all-reduce when symmetric memory is unavailable)" if world > 1 else ""),
           "l2": "inputs larger than L2: every step streams the whole p/g/m/v arena through Adam "
                 "(hash: 489 MB/step) plus per-batch activations (SIREN: > 1 GB/layer)"}
    cfg["allocator_priming_steps_before_warmup"] = PRIMING_STEPS
    cfg["batch_order"] = (os.environ.get("MRI_BATCH_ORDER", "axis0") + ": every batch arranged with the axis-0 voxel index "
                          "fastest (same sets; the loss is order-invariant)")
    if cpu_sample:
        cfg["cpu_sample_coords_per_step"] = cpu_sample
    return cfg


# -------------------------------------------------------------------------------------------------- B200 arm
def synthetic_volume(shape, dev, seed=1337):
    """Deterministic smooth + noise field on the device (config 4/5): sum of separable sinusoids + 1% hash noise."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    out = torch.zeros(shape, device=dev, dtype=torch.float32)
    for k in range(4):
        term = torch.ones(shape, device=dev, dtype=torch.float32)
        for d, s in enumerate(shape):
            f = 1 + ((seed + 3 * k + d) % 5)
            ph = 0.37 * (k + 1) * (d + 1)
            view = [1] * len(shape)
            view[d] = s
            term = term * torch.sin(torch.linspace(0, 1, s, device=dev) * (6.2831853 * f) + ph).reshape(view)
        out += term
        del term
    out += 0.04 * torch.rand(shape, device=dev, generator=g)
    out = (out - out.min()) / (out.max() - out.min())
    return out.flatten()


def build_workload(name, dev):
    from mri_interpolation_b200 import models, nifti
    from mri_interpolation_b200 import functional as Fn
    torch.manual_seed(1337)
    info = {}
    if name in ("ankle_hash", "synthetic_hash"):
        model = models.HashMLP(dim_in=4, dim_hidden=64, dim_out=1, n_layers=2, lr=5e-3, batch_norm=False, **G4).to(dev)
        info["train_flops_per_coord"] = None
    elif name == "siren_wide":
        model = models.SirenNet(dim_in=3, dim_hidden=1024, dim_out=1, n_layers=8, w0=30.0, w0_initial=30.0, lr=1e-4).to(dev)
        info["fwd_flops_per_coord"], info["train_flops_per_coord"] = 14.688e6, 44.06e6
    else:
        model = models.SirenNet(dim_in=4, dim_hidden=256, dim_out=1, n_layers=5, w0=30.0, w0_initial=30.0, lr=1e-4).to(dev)
        info["fwd_flops_per_coord"], info["train_flops_per_coord"] = 0.527e6, 1.58e6
    norm_siren = name.startswith("siren")
    if name in ("ankle_hash", "siren_ankle"):
        vol = nifti.load(SAMPLE).get_fdata(np.float32)
        shape = vol.shape
        pix = torch.from_numpy(vol).flatten()
        pix = ((pix - pix.min()) / (pix.max() - pix.min())).to(dev)
    elif name == "synthetic_hash":
        shape = (256, 256, 256, 32)
        pix = synthetic_volume(shape, dev)
    else:
        shape = (512, 512, 512)
        pix = synthetic_volume(shape, dev)
    if norm_siren:
        pix = pix * 2 - 1
    sampler = Fn.VoxelSampler(pix, shape, norm_siren=norm_siren)
    info["shape"] = shape
    info["data"] = ("sample_ankle_dyn_mri.nii.gz (bundled reference sample volume)" if "ankle" in name
                    else f"synthetic {'x'.join(map(str, shape))} volume generated on device") + ", random-init weights (seed 1337)"
    return model, sampler, info


def draw_index_ring(name, sampler, info, n, ring, dev, rank, world):
    """(ring, n) voxel indices drawn before the timed region + what drawing one batch costs (ms, amortised)."""
    from mri_interpolation_b200 import functional as Fn
    from mri_interpolation_b200.datamodules import ShuffledEpochs
    order = os.environ.get("MRI_BATCH_ORDER", "axis0")
    shape = info["shape"]
    if "ankle" in name:
        # shuffled epochs of the sample volume, as the reference's DataLoader(shuffle=True): a permutation cut into batches
        # (data-parallel ranks take disjoint strided shares), locality-ordered inside each batch
        epochs = ShuffledEpochs(sampler.total, n, dev, seed=1337, rank=rank, world_size=world,
                                grid_shape=shape if order != "none" else None)
        full = epochs.local_count() // n  # whole batches per epoch
        if full == 0:
            raise SystemExit(f"batch of 2^{int(np.log2(n))} per GPU exceeds this rank's share of the volume")
        parts, got, n_epochs = [], 0, 0
        epochs.epoch()  # untimed: the first call pays torch's sort / randperm workspace allocations
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        while got < ring:
            parts.append(epochs.epoch()[: full * n].reshape(full, n))
            got += full
            n_epochs += 1
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / (n_epochs * full)
        index = torch.cat(parts)[:ring].contiguous()
        how = (f"ShuffledEpochs: {full} whole batches per epoch of this rank's {epochs.local_count()} voxels, locality order "
               f"{order}; ms = epoch generation / batches per epoch")
    else:
        gen = torch.Generator(device=dev)
        gen.manual_seed(1337 + rank)

        def draw():
            i = torch.randint(0, sampler.total, (n,), device=dev, generator=gen)
            return Fn.locality_sort(i, shape, block=1) if order != "none" else i

        index = torch.stack([draw() for _ in range(ring)])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(4):
            draw()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 4 * 1e3
        how = f"i.i.d. torch.randint per batch (SURVEY 8d config 4/5), locality order {order} by one argsort per batch"
    return index, ms, how


def traffic_probe(args, batch_log2):
    """roofline.traffic measured on THIS box in THIS run: a nested `ncu --metrics ...` run of this script (--probe-steps:
    the same workload and batch size, a handful of plain training steps) captures one steady-state launch of each fused
    kernel and returns {kernel: {metric: value}}.  Byte / sector counters only - no timing is taken under the profiler.
    Returns (None, reason) when ncu is absent or the counters are not accessible."""
    import csv
    import io
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.isfile(ncu):
        return None, "ncu not found"
    cmd = [ncu, "--metrics", ",".join(PROBE_METRICS), "--clock-control", "none", "--print-units", "base", "-k",
           "regex:hashdecoder_mma_(fwd|bwd)_kernel", "-s", "8", "-c", "2", "--csv", sys.executable, os.path.abspath(__file__),
           "--probe-steps", "8", "--workload", args.workload, "--batch-log2", str(batch_log2)]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
    try:
        run = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    except Exception as e:  # noqa: BLE001
        return None, f"ncu probe failed: {type(e).__name__}"
    rows = list(csv.reader(io.StringIO(run.stdout[run.stdout.find('"ID"'):]))) if '"ID"' in run.stdout else []
    if run.returncode != 0 or len(rows) < 2:
        return None, f"ncu probe rc={run.returncode}: {(run.stdout + run.stderr).strip()[-200:]}"
    col = {h: i for i, h in enumerate(rows[0])}
    out = {}
    for r in rows[1:]:
        if len(r) <= col["Metric Value"]:
            continue
        kn = r[col["Kernel Name"]]
        # the one-kernel training step is the backward template with STEP = true (last template argument)
        kname = ("hashmlp_step" if "(bool)1, (bool)1>" in kn else
                 "hashdecoder_bwd" if "hashdecoder_mma_bwd" in kn else "hashdecoder_fwd")
        try:
            out.setdefault(kname, {})[r[col["Metric Name"]]] = float(r[col["Metric Value"]].replace(",", ""))
        except ValueError:
            pass
    return (out, "ncu, nested run of this command on this box") if out else (None, "ncu probe returned no metrics")


def probe_main(args):
    """Child of traffic_probe: the headline workload's plain training steps, nothing timed, nothing printed."""
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    model, sampler, info = build_workload(args.workload, dev)
    opt = model.configure_optimizers()
    n = 1 << args.batch_log2
    index, _, _ = draw_index_ring(args.workload, sampler, info, n, 4, dev, 0, 1)
    from mri_interpolation_b200.pl_compat import training_step_and_backward
    for i in range(args.probe_steps):
        x, y = sampler.batch(index[i % 4])
        training_step_and_backward(model, (x, y), i)
        opt.step()
        opt.zero_grad()
    torch.cuda.synchronize()


def fit_throughput(dev, batch_log2, epochs=8):
    """The launcher's own path, sampler and Python loop included: MriDataModule (the sample volume as device tensors,
    shuffled epochs drawn on the fly) -> pl_compat.Trainer.fit -> LightningModule.training_step / FusedAdam, wall clock
    around fit() after one untimed epoch.  coords/s = batches x batch size / seconds."""
    from mri_interpolation_b200 import config as cfgmod, datamodules, models
    from mri_interpolation_b200.pl_compat import pl
    cfg = cfgmod.HashConfig()
    cfg.image_path, cfg.batch_size = SAMPLE, 1 << batch_log2
    dm = datamodules.MriDataModule(config=cfg, device=dev)
    dm.prepare_data()
    loader = dm.train_dataloader()
    torch.manual_seed(1337)
    model = models.HashMLP(dim_in=4, dim_hidden=64, dim_out=1, n_layers=2, lr=5e-3, batch_norm=False, **G4)
    steps = epochs * len(loader)
    full = epochs * loader.epochs.local_count()

    def measure(graph):
        trainer = pl.Trainer(accelerator="gpu", max_epochs=1, precision=32, enable_checkpointing=False, logger=False, cuda_graph=graph)
        trainer.fit(model, loader)  # untimed: allocator, first-call workspaces
        torch.cuda.synchronize()
        runs = []
        for _ in range(3):  # wall-clock timing of a few hundred ms is noisy on these shared hosts: median of three
            trainer = pl.Trainer(accelerator="gpu", max_epochs=epochs, precision=32, enable_checkpointing=False, logger=False,
                                 cuda_graph=graph)
            t0 = time.perf_counter()
            trainer.fit(model, loader)
            torch.cuda.synchronize()
            runs.append(time.perf_counter() - t0)
        dt = float(np.median(runs))
        return {"value": full / dt, "unit": "coords/s", "seconds": dt, "seconds_per_run": [round(r, 4) for r in runs],
                "ms_per_step": dt / steps * 1e3}

    out = measure(False)
    out.update({"epochs": epochs, "steps": steps,
                "note": "Trainer.fit over MriDataModule.train_dataloader() on the sample volume: shuffled-epoch sampling, "
                        "batch gather, Python loop, logging, optimiser construction and the ragged last batch of every epoch "
                        "all inside the clock; median of three runs (batches come from mri_gather_voxels, training_step + "
                        "backward run as direct kernel calls - HashMLP.fused_training_step)"})
    del model, loader, dm
    torch.cuda.empty_cache()
    return out


def red_rate_peak(dev):
    """Peak L2 sector-reduction rate of this GPU (csrc/probe.cu), G sector-ops/s: {spread: 32 sectors per warp
    instruction, paired: adjacent lanes share a sector like the pair-lane scatter}."""
    import ctypes
    from mri_interpolation_b200 import _lib
    table = torch.zeros(16 << 20, device=dev)  # 64 MB: L2-resident like the 61 MB gradient arena
    ops = ctypes.c_int64(0)
    out = {}
    for name, lanes in (("spread", 1), ("paired", 2)):
        best = None
        for rep in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.call("mri_probe_red_rate", table.data_ptr(), table.numel(), 64, lanes, ctypes.byref(ops), _lib.stream())
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b)
            best = ms if best is None or (rep > 0 and ms < best) else best
        out[name] = ops.value / (best * 1e-3) / 1e9
    del table
    return out


def run_leg(args, name, batch_log2, steps, warmup, dev, rank, local_rank, world, *, e2e=True, infer=True, kernels=True,
            settle_s=0.0):
    """One workload: timed training steps (+ sustained window with clocks), e2e, inference sweep, isolated kernels."""
    import torch.distributed as dist
    from mri_interpolation_b200 import _lib, sweep
    from mri_interpolation_b200.datamodules import PrefetchLoader

    model, sampler, info = build_workload(name, dev)
    is_hash = name.endswith("hash")
    opt = model.configure_optimizers()
    # bucketed/overlapped all-reduce is available but measured SLOWER at W=2 (1.608 vs 1.561 ms/step): the scatter
    # backward and NCCL's copy kernels both saturate the L2, so the default stays one exchange after backward
    if is_hash and world > 1 and os.environ.get("MRI_DP_OVERLAP") == "1":
        opt.enable_overlap(model.encoder, n_groups=4)
    n = 1 << batch_log2
    dim = len(info["shape"])
    ring = 32
    index, sampler_ms, sampler_how = draw_index_ring(name, sampler, info, n, ring, dev, rank, world)

    from mri_interpolation_b200.pl_compat import training_step_and_backward

    def step(i):
        # exactly what pl_compat.Trainer.fit runs per batch: training_step + backward (HashMLP under the MSE loss: ONE kernel
        # for both, csrc/hashdecoder_step.cu; MRI_FUSED_STEP=0 restores forward kernel + MSE kernel + backward kernel)
        x, y = sampler.batch(index[i % ring])
        loss = training_step_and_backward(model, (x, y), i)
        opt.step()
        opt.zero_grad()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(first, count):
        """`count` steps with the host's run-ahead bounded to two steps (a saturated launch queue starves NCCL's progress
        thread and made multi-GPU SIREN steps ~25% slower); returns (device ms, last loss)."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        inflight = []
        ev0.record()
        loss = None
        for i in range(first, first + count):
            loss = step(i)
            ev = torch.cuda.Event()
            ev.record()
            inflight.append(ev)
            if len(inflight) > 2:
                inflight.pop(0).synchronize()
        ev1.record()
        barrier()
        return ev0.elapsed_time(ev1), loss

    def max_over_ranks(v):
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()  # nvidia-smi needs a few hundred ms to come up: started before the untimed steps, marked below
    for i in range(PRIMING_STEPS):  # untimed: lets torch's caching allocator reach its steady state (no cudaMalloc later)
        step(i)
    for i in range(warmup):
        step(i)
    if settle_s > 0:
        # secondary legs (tens of steps): run untimed until the board has been under this load for ~settle_s, so the
        # power-cap clock transient of the first second (SIREN legs draw ~1 kW) is not what the short timed region sees.
        # The step count is agreed over the ranks first (the steps are collective).
        ms5, _ = run_steps(0, 5)
        n_settle = int(max_over_ranks(settle_s * 1e3 / max(ms5 / 5, 1e-3))) + 1
        run_steps(0, min(n_settle, 2000))
    barrier()
    clocks.mark()
    launches0 = _lib.launch_count
    ms, loss = run_steps(warmup, steps)
    launches = _lib.launch_count - launches0
    ms_step = max_over_ranks(ms) / steps
    final_loss = float(loss.detach())
    # the same steps again for >= SUSTAIN_SECONDS (same count on every rank: the steps are collective), so that the clock
    # samples describe the hardware state under this very load even when --steps is small
    more = max(20, int(SUSTAIN_SECONDS * 1e3 / max(ms_step, 1e-3)) + 1)
    ms2, _ = run_steps(warmup + steps, more)
    sustained_ms_step = max_over_ranks(ms2) / more
    clk = clocks.stop() if rank == 0 else None
    if clk is not None:
        clk["window"] = f"timed region ({steps} steps) + sustained window ({more} identical steps, {ms2 / 1e3:.2f} s)"
    value = n * world / (ms_step * 1e-3)

    # optimiser step as seen inside the training step (includes the gradient exchange at N > 1)
    opt_ms = []
    for i in range(12):
        x, y = sampler.batch(index[i % ring])
        training_step_and_backward(model, (x, y), i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); opt.step(); b.record()
        opt.zero_grad()
        barrier()
        opt_ms.append(a.elapsed_time(b))
    opt_step_ms = max_over_ranks(float(np.median(opt_ms[2:])))
    exchange_parts = None
    if world > 1 and getattr(opt, "sharded", False):
        # where the exchange time goes: opening barrier (includes waiting for the slowest rank's backward), the fused
        # reduce-scatter + Adam + all-gather kernel, closing barrier, local gradient clear (device times, median, max over ranks)
        opt.time_exchange, opt.exchange_times = True, []
        for i in range(12):
            x, y = sampler.batch(index[i % ring])
            training_step_and_backward(model, (x, y), i)
            opt.step()
            opt.zero_grad()
        opt.time_exchange = False
        t = np.median(np.array(opt.exchange_times[2:]), axis=0)
        exchange_parts = {k: max_over_ranks(float(v)) for k, v in zip(("open_barrier_ms", "kernel_ms", "close_barrier_ms", "clear_ms"), t)}

    # ---- e2e: host batches through the public API (PrefetchLoader + LightningModule.training_step + FusedAdam)
    # every step: H2D copy of that step's batch from pinned host memory (overlapped with the previous step's
    # compute on a side stream) and a D2H read of that step's loss (async copy, consumed one step later).
    e2e_line = None
    if e2e:
        host_ring = []
        for r in range(4):
            xb, yb = sampler.batch(index[r])
            host_ring.append((xb.cpu().pin_memory(), yb.cpu().pin_memory()))
        e2e_steps = max(10, min(steps, 100))

        class _HostBatches:
            def __len__(self):
                return e2e_steps + 3

            def __iter__(self):
                for i in range(len(self)):
                    yield host_ring[i % len(host_ring)]

        loss_host = torch.zeros(e2e_steps + 3, dtype=torch.float32).pin_memory()
        loss_events = [torch.cuda.Event() for _ in range(e2e_steps + 3)]
        losses = []
        t0 = None
        for i, (xb, yb) in enumerate(PrefetchLoader(_HostBatches(), dev)):
            if i == 3:  # 3 untimed warm-up steps
                barrier()
                t0 = time.perf_counter()
            l = model.training_step((xb, yb), i)
            l.backward()
            opt.step()
            opt.zero_grad()
            loss_host[i].copy_(l.detach(), non_blocking=True)
            loss_events[i].record()
            if i >= 1:
                loss_events[i - 1].synchronize()
                losses.append(float(loss_host[i - 1]))
        barrier()
        e2e_dt = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
        e2e_line = {"value": n * world / e2e_dt, "unit": "coords/s", "h2d_bytes_per_step": n * (dim + 1) * 4,
                    "d2h_bytes_per_step": 4, "steps": e2e_steps}

    # ---- inference: every rank sweeps its slab of the query volume (no communication), max over ranks
    infer_line = None
    if infer:
        if is_hash:
            sweep_shape = SWEEP_SHAPE if name == "ankle_hash" else (256, 256, 256, 4)
        else:
            sweep_shape = (352, 352, 6, 29) if name == "siren_ankle" else (256, 256, 128)
        total_vox = int(np.prod(sweep_shape))
        ns = not is_hash
        # everything that does not depend on the query (axis vectors, eval-mode BatchNorm folding, output buffers) is
        # prepared once: the timed window holds the sweep kernels only
        sweeper = sweep.SlabSweeper(model, sweep_shape, norm_siren=ns, rank=rank, world_size=world)
        for _ in range(2):
            sweeper.run()
        barrier()
        reps, per_rep = 5, []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            sweeper.run()
            b.record()
            barrier()
            per_rep.append(a.elapsed_time(b))
        # median of the per-sweep device times: single sweeps occasionally stall for tens of ms on these shared boxes
        sw_ms = max_over_ranks(float(np.median(per_rep)))
        # end to end: the slab lands in pinned host memory, chunk by chunk, the copy of chunk i overlapping chunk i+1
        sweeper.run_to_host()
        barrier()
        e2e_times = []
        for _ in range(3):
            t1 = time.perf_counter()
            sweeper.run_to_host()
            e2e_times.append(time.perf_counter() - t1)
            barrier()
        e2e_s = max_over_ranks(float(np.median(e2e_times)))
        infer_line = {"metric": "infer_voxels_per_s", "value": total_vox / (sw_ms * 1e-3), "unit": "voxels/s", "n_gpus": world,
                      "workload": f"dense sweep of {sweep_shape} ({total_vox} voxels), contiguous slab per GPU, "
                                  + ("fused hash+decoder kernel, axis-0-fastest walk" if is_hash else "coordinate synthesis + tensor-core SIREN"),
                      "ms": sw_ms, "ms_per_sweep": [round(v, 3) for v in per_rep],
                      "e2e": {"value": total_vox / e2e_s, "unit": "voxels/s", "h2d_bytes_per_step": 0,
                              "d2h_bytes_per_step": 4 * total_vox // world,
                              "note": "slab copied to pinned host memory in chunks on a side stream while the next chunk computes"}}
        del sweeper

    # ---- isolated kernel timings for the roofline (rank 0, L2 flushed between launches)
    roof, kern = None, None
    if kernels:
        if is_hash:
            opt.disable_overlap(model.encoder)
        opt.data_parallel = False  # the isolated-kernel section below runs on rank 0 only: no collectives in it
    if kernels and rank == 0:
        pk, peak_src = peaks()
        hbm_peak = float(pk["hbm_gbs"])
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def timed(fn, reps=10):
            times = []
            for _ in range(3):
                fn()
            for _ in range(reps):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record()
                torch.cuda.synchronize()
                times.append(a.elapsed_time(b))
            return float(np.mean(times))

        if is_hash:
            enc = model.encoder
            x, y = sampler.batch(index[0])
            go = torch.randn(n, 32, device=dev)
            with torch.no_grad():
                fwd_ms = timed(lambda: enc(x))
            enc_out = enc(x)
            bwd_ms = timed(lambda: torch.autograd.backward(enc_out, go, retain_graph=True))
            # the whole-model forward / backward (ONE kernel each when the decoder is fused with the gather / scatter)
            fused_ms = fused_fwd_ms = None
            if getattr(model, "fuse_backward", False):
                pred = model(x)
                gy = torch.randn_like(pred) * 1e-3
                if pred.grad_fn is not None and type(pred.grad_fn).__name__.startswith("HashDecoderFn"):
                    fused_ms = timed(lambda: torch.autograd.backward(pred, gy, retain_graph=True))
                    fused_fwd_ms = timed(lambda: model(x))  # training forward: gather + decoder, enc and pre2 written
                del pred
            # the one-kernel training step (gather + decoder + MSE + decoder backward + scatter), when the model has it
            step_ms = None
            if getattr(model, "fuse_step", False) and model.fused_training_step((x, y), 0) is not None:  # opt-in one-kernel step
                step_ms = timed(lambda: model.fused_training_step((x, y), 0))
            opt.arena.grad.zero_()
            # the single-GPU Adam kernel on scratch arenas of the model's size (the model itself is not stepped here)
            scratch = [torch.zeros(opt.arena.numel, device=dev) for _ in range(4)]
            adam_ms = timed(lambda: _lib.call("mri_adam_step", scratch[0].data_ptr(), scratch[1].data_ptr(), scratch[2].data_ptr(),
                                              scratch[3].data_ptr(), opt.arena.numel, 1, 5e-3, 0.9, 0.999, 1e-8, 0.0, 1.0, 1,
                                              _lib.stream()))
            del scratch
            hash_bytes = HASH_BYTES_PER_COORD * n
            adam_bytes = 32 * opt.arena.numel  # p,g,m,v read; p,m,v,g written (fused gradient clear)
            kern = {
                "hashgrid_fwd": {"ms": fwd_ms, "GBps_algorithmic": hash_bytes / fwd_ms / 1e6, "frac": hash_bytes / fwd_ms / 1e6 / hbm_peak},
                "hashgrid_bwd": {"ms": bwd_ms, "GBps_algorithmic": hash_bytes / bwd_ms / 1e6, "frac": hash_bytes / bwd_ms / 1e6 / hbm_peak},
                "adam_step": {"ms": adam_ms, "GBps_algorithmic": adam_bytes / adam_ms / 1e6, "frac": adam_bytes / adam_ms / 1e6 / hbm_peak},
            }
            top = "hashgrid_bwd" if bwd_ms >= fwd_ms else "hashgrid_fwd"
            traffic = None
            if fused_ms is not None:
                # fused decoder-backward + scatter: coords + enc + dy read, L*2^D*F*4 B reduced into the tables
                fused_bytes = (HASH_BYTES_PER_COORD + 4) * n
                kern["hashdecoder_bwd"] = {"ms": fused_ms, "GBps_algorithmic": fused_bytes / fused_ms / 1e6,
                                           "frac": fused_bytes / fused_ms / 1e6 / hbm_peak}
                kern["hashdecoder_fwd"] = {"ms": fused_fwd_ms, "GBps_algorithmic": (fused_bytes + 4 * n) / fused_fwd_ms / 1e6,
                                           "frac": (fused_bytes + 4 * n) / fused_fwd_ms / 1e6 / hbm_peak}
                top = "hashdecoder_bwd" if fused_ms >= fused_fwd_ms else "hashdecoder_fwd"
                hash_bytes = fused_bytes if top == "hashdecoder_bwd" else fused_bytes + 4 * n
                # DRAM bytes and L2 reduction sectors per launch: measured now by a nested ncu run (headline leg only),
                # else the committed capture (2^19 ankle-volume coords; other sizes / volumes have none)
                probe, traffic_src = (None, "probe skipped")
                if name == args.workload and not args.no_traffic_probe and world == 1:
                    probe, traffic_src = traffic_probe(args, batch_log2)
                red_sectors = None
                if probe and top in probe and "dram__bytes_read.sum" in probe[top]:
                    traffic = probe[top]["dram__bytes_read.sum"] + probe[top]["dram__bytes_write.sum"]
                    red_sectors = probe.get("hashdecoder_bwd", {}).get("lts__t_sectors_srcunit_tex_op_red.sum")
                elif name == "ankle_hash" and batch_log2 == 19:
                    traffic, traffic_src = NCU_DRAM_BYTES[top], NCU_SOURCE + f" (live probe: {traffic_src})"
                    red_sectors = NCU_RED_SECTORS["hashdecoder_bwd"]
                else:
                    traffic_src = None
            if step_ms is not None and fused_ms is not None:
                # ONE kernel for forward + loss + backward: coords + target read, L*2^D*F*4 B gathered and as many reduced
                fused_step_bytes = (4 * 4 + 4 + 2 * 16 * 16 * 2 * 4) * n
                kern["hashmlp_step"] = {"ms": step_ms, "GBps_algorithmic": fused_step_bytes / step_ms / 1e6,
                                        "frac": fused_step_bytes / step_ms / 1e6 / hbm_peak}
                top, hash_bytes = "hashmlp_step", fused_step_bytes
                if probe and top in probe and "dram__bytes_read.sum" in probe[top]:
                    traffic = probe[top]["dram__bytes_read.sum"] + probe[top]["dram__bytes_write.sum"]
                    red_sectors = probe[top].get("lts__t_sectors_srcunit_tex_op_red.sum")
                else:
                    traffic, traffic_src = None, f"none for the one-kernel step (live probe: {traffic_src})"
                step_bytes = fused_step_bytes + adam_bytes
            else:
                step_bytes = (2 * HASH_BYTES_PER_COORD + 8) * n + adam_bytes  # fused fwd + fused bwd + Adam, algorithmic
            kern["whole_step"] = {"ms": ms_step, "algorithmic_GB": step_bytes / 1e9, "GBps_algorithmic": step_bytes / ms_step / 1e6,
                                  "frac": step_bytes / ms_step / 1e6 / hbm_peak}
            roof = {"kernel": top, "bound": "hbm", "achieved": kern[top]["GBps_algorithmic"], "peak": hbm_peak, "unit": "GB/s",
                    "frac": kern[top]["frac"], "traffic": traffic, "traffic_source": traffic_src if fused_ms is not None else None,
                    "peak_source": peak_src + " hbm_gbs", "algorithmic_bytes_per_launch": hash_bytes,
                    "note": "one-kernel training step: 4116 B/coord = 4D coords + target + L*2^D*F*4 gathered + L*F*2^D*4 reduced "
                            "(the encoding never goes to memory); two-kernel path: 2192 B/coord per kernel (4D + L*2^D*F*4 "
                            "gathered-or-reduced + L*F*4 encoding, +4 B/coord dy for the backward) and 8.4 kMAC/coord of decoder math; "
                            "traffic = ncu dram read+write per launch: the 61 MB of tables stay in the 126 MB L2, so DRAM moves "
                            "less than the algorithmic bytes - the kernels are bound by the L2 sector / L2 atomic rates "
                            "(DESIGN.md 3), not HBM"}
            if fused_ms is not None and red_sectors:
                # the resource that actually binds the backward: 32-byte sector reductions retired by the L2 atomic units
                peak = red_rate_peak(dev)
                bind_ms = step_ms if top == "hashmlp_step" else fused_ms
                ach = red_sectors / (bind_ms * 1e-3) / 1e9
                roof["binding"] = {"resource": "L2 atomic units: 32-byte sector reductions (red.global.add.v2.f32)",
                                   "kernel": top if top == "hashmlp_step" else "hashdecoder_bwd", "sector_ops_per_launch": red_sectors, "achieved": ach,
                                   "peak": max(peak["paired"], peak["spread"]), "peak_paired": peak["paired"], "peak_spread": peak["spread"],
                                   "unit": "G sector-ops/s", "frac": ach / max(peak["paired"], peak["spread"]),
                                   "note": "frac ~ 1 (it can exceed 1 by a few per cent: the probe's two synthetic patterns are a measurement "
                                           "of the ceiling, not a bound on it - scripts/lsu_microbench.cu puts it at 92 sector operations per "
                                           "clock = 181 G/s)",
                                   "peak_source": "measured in this run by mri_probe_red_rate (csrc/probe.cu): every warp instruction "
                                                  "reduces into 16 (paired, the scatter's own pattern) / 32 (spread) distinct sectors "
                                                  "of an L2-resident 64 MB table, 4 CTAs of 128 threads per SM",
                                   "sector_source": traffic_src}
        else:
            from mri_interpolation_b200 import tc
            from mri_interpolation_b200._lib import ACT_SINE
            h = model.dim_hidden
            a_f = torch.rand(n, h, device=dev) * 2 - 1
            w_f = model.layers[1].weight.detach()
            a_hi, a_lo = tc.split(a_f)
            w_hi, w_lo = tc.split(w_f)
            bias = model.layers[1].bias.detach()
            ms_l = timed(lambda: tc.layer(a_hi, a_lo, w_hi, w_lo, bias, ACT_SINE, 30.0, passes=3, want_planes=True, want_aux=True), reps=5)
            gw = torch.zeros(h, h, device=dev)
            ms_w = timed(lambda: tc.wgrad(a_hi, a_lo, a_hi, a_lo, gw, None, passes=3), reps=5)
            tf_burst, tf_sust = float(pk["bf16_tflops"]), float(pk.get("bf16_tflops_sustained", pk["bf16_tflops"]))
            issued = 3 * 2.0 * n * h * h
            kern = {"siren_tc_layer_fwd": {"ms": ms_l, "issued_bf16_TFLOPs": issued / ms_l / 1e9, "frac": issued / ms_l / 1e9 / tf_burst},
                    "siren_tc_wgrad": {"ms": ms_w, "issued_bf16_TFLOPs": issued / ms_w / 1e9, "frac": issued / ms_w / 1e9 / tf_burst}}
            # a hidden layer moves 12 B per activation element (bf16 hi/lo planes in, planes + fp32 w0 cos(.) out) for
            # 6 H issued flops: tensor-bound for wide layers, HBM-bound below H ~ 500 - the line reports the binding one
            layer_bytes = 12.0 * n * h + 8.0 * h * h
            hbm_gbs = layer_bytes / ms_l / 1e6
            kern["siren_tc_layer_fwd"]["GBps_algorithmic"] = hbm_gbs
            kern["siren_tc_layer_fwd"]["frac_hbm"] = hbm_gbs / hbm_peak
            tensor_roof = {"kernel": "siren_tc_layer_kernel (hidden layer forward, bf16x3 split precision, sine epilogue)",
                           "bound": "tensor", "achieved": kern["siren_tc_layer_fwd"]["issued_bf16_TFLOPs"], "peak": tf_burst,
                           "unit": "TFLOP/s", "frac": kern["siren_tc_layer_fwd"]["frac"], "traffic": None,
                           "peak_source": peak_src + " bf16_tflops (burst: the kernel is timed alone)",
                           "algorithmic_flops_per_launch": 2.0 * n * h * h,
                           "note": "issued = 3 tcgen05.mma passes x 2 n H^2 (A_lo*B_hi + A_hi*B_lo + A_hi*B_hi); the fp32-equivalent "
                                   "(algorithmic) rate is one third of it"}
            if hbm_gbs / hbm_peak > tensor_roof["frac"]:
                roof = {"kernel": tensor_roof["kernel"], "bound": "hbm", "achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s",
                        "frac": hbm_gbs / hbm_peak, "traffic": None, "peak_source": peak_src + " hbm_gbs",
                        "algorithmic_bytes_per_launch": layer_bytes,
                        "note": f"H = {h}: 12 B per activation element (bf16 hi/lo planes read, planes + fp32 w0 cos written) + the "
                                f"weight planes; arithmetic intensity {issued / layer_bytes:.0f} issued flop/B is below the machine "
                                f"balance, so the layer is HBM-bound (tensor-pipe figure kept under `tensor`)",
                        "tensor": {k: tensor_roof[k] for k in ("achieved", "peak", "unit", "frac")}}
            else:
                roof = tensor_roof
                roof["hbm"] = {"achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_gbs / hbm_peak}
            if info.get("train_flops_per_coord"):
                alg = info["train_flops_per_coord"] * n / (ms_step * 1e-3) / 1e12
                kern["whole_step"] = {"ms": ms_step, "algorithmic_TFLOPs": alg, "issued_bf16_TFLOPs": 3 * alg,
                                      "frac_of_sustained_peak": 3 * alg / tf_sust,
                                      "note": "in-step figure against bf16_tflops_sustained (a long step runs under the power cap)"}
        del flush

    line = None
    if rank == 0:
        line = {
            "metric": "train_coords_per_s", "value": value, "unit": "coords/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if is_hash else "f32 via bf16x3 split on tcgen05 (fp32 accumulate)",
            "data": info["data"], "config": workload_config(args, name, batch_log2, world), "clocks": clk,
            "sustained": {"steps": more, "ms_per_step": sustained_ms_step, "value": n * world / (sustained_ms_step * 1e-3)},
            "e2e": e2e_line, "gpu_launches": launches, "final_loss": final_loss,
            "sampler_ms_per_batch": sampler_ms, "sampler": sampler_how,
            "optimizer_step_ms": opt_step_ms, "exchange_parts": exchange_parts,
            "roofline": roof, "kernels": kern, "infer": infer_line,
        }
        if settle_s > 0:
            line["config"]["untimed_settle_s_after_warmup"] = settle_s
        if kern and "adam_step" in kern:
            # what the gradient exchange adds to the optimiser step: in-step optimiser time minus the single-GPU Adam kernel
            line["exchange_exposed_ms"] = max(0.0, opt_step_ms - kern["adam_step"]["ms"]) if world > 1 else 0.0
    del model, opt, sampler, index
    torch.cuda.empty_cache()
    return line


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    if args.probe_steps > 0:
        probe_main(args)
        return

    from mri_interpolation_b200 import distributed

    # the CPU baseline runs BEFORE the process group exists (rank 0 of a single-GPU run only: under torchrun the other
    # ranks would spin in a barrier on the same host cores and contaminate it - round-1 finding)
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    cpu = None
    if world_env == 1 and not args.no_cpu_baseline and args.workload == "ankle_hash":
        v, dt, ncpu, kind = time_reference(args.cpu_batch_log2, 2, 1)
        cpu = {"value": v, "unit": "coords/s", "cores": torch.get_num_threads(), "kind": kind,
               "sample": cpu_sample_text(kind, 2, args.cpu_batch_log2, dt)}

    # NCCL_DEBUG=VERSION (set in this image) makes NCCL print its version banner on STDOUT, in front of the one JSON line
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    rank, local_rank, world = distributed.init_from_env("nccl")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    import torch.distributed as dist

    line = run_leg(args, args.workload, args.batch_log2, args.steps, args.warmup, dev, rank, local_rank, world,
                   e2e=not args.no_e2e, infer=not args.no_infer, kernels=True)
    fit = None
    if rank == 0 and world == 1 and args.workload == "ankle_hash" and not args.no_e2e:
        # right after the headline leg, before the secondary legs heat the board and churn the allocator
        try:
            fit = fit_throughput(dev, args.batch_log2)
        except Exception as e:  # noqa: BLE001 - a reported extra, never a reason to lose the bench line
            fit = {"error": f"{type(e).__name__}: {e}"}
    extra = {}
    if not args.no_workloads and args.workload == "ankle_hash":
        for name in EXTRA_LEGS:
            steps = 20 if name == "siren_wide" else 40
            leg = run_leg(args, name, WORKLOADS[name][1], steps, 3, dev, rank, local_rank, world,
                          e2e=False, infer=(name != "synthetic_hash"), kernels=True, settle_s=0.6)
            if rank == 0:
                extra[name] = {k: leg[k] for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "dtype", "data",
                                                   "config", "clocks", "sustained", "gpu_launches", "final_loss", "roofline", "kernels",
                                                   "infer", "optimizer_step_ms", "sampler_ms_per_batch")}
    if rank == 0:
        line["cpu_baseline"] = cpu
        if fit is not None:
            line["fit"] = fit
        if extra:
            line["workloads"] = extra
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
