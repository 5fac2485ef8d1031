#!/usr/bin/env python
"""Benchmark of the coordinate-network hot path (BASELINE.json metric: training coords/s, plus inference
voxels/s, with % of roofline and the reference's CPU path timed beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch-log2 B]

Workload (configs[1] of BASELINE.json, named in `config.workload`): hash-grid encoder G4
(config/hash_config.json: L=16, F=2, T=2^19, base 16, x1.4 -> finest 2489; 15 279 648 table params)
+ 2-layer GELU decoder (64 hidden), fitted to the sample ankle volume (352x352x6x15, x,y,z,t coords),
fp32, Adam lr 5e-3, 2^B coordinates per step PER GPU (weak scaling; N>1 exchanges the flat gradient arena once
per step inside the sharded Adam kernel).  One step = 5 kernels: sample voxel indices -> synthesise coords /
gather intensities -> hash encode + decoder (one kernel) -> MSE -> decoder backward + hash scatter (one
kernel) -> fused Adam [with the reduce-scatter / all-gather over NVLink at N>1].

value : device-resident inputs, CUDA-event timing, max over ranks.
e2e   : same step through the public LightningModule API with HOST (pinned) batches: H2D copy of the
        batch and D2H read of the loss inside the timed region.
roofline : dominant kernel (decoder backward fused with the hash-grid scatter) timed alone with CUDA events, L2
        flushed between launches; algorithmic bytes / duration against MEASURED_PEAKS.json's HBM copy bandwidth.
cpu_baseline : the oracle port of the reference's PyTorch path on this box's host cores (bounded sample).
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
# dram__bytes_read.sum + dram__bytes_write.sum per launch of hashdecoder_mma_bwd_kernel at 2^19 coords (ncu --set full,
# profiles/r01_ncu_full_fused_hashdecoder.csv): 137.9 MB read + 256.8 MB written
FUSED_BWD_DRAM_BYTES = 394.6e6


WORKLOADS = {
    # name: (description, default log2 batch per GPU)
    "ankle_hash": ("hash-grid G4 (L16 F2 T2^19 base16 finest2489, 15.28M table params) + 2x64 GELU decoder fitted to "
                   "sample_ankle_dyn_mri.nii.gz (352x352x6x15, xyzt coords), Adam lr 5e-3", 19),
    "synthetic_hash": ("config 4: hash-grid G4 + 2x64 GELU decoder on a synthetic 256^3 x 32-frame volume (536.9M voxels, "
                       "sum of separable sinusoids + noise, generated on device), Adam lr 5e-3", 19),
    "siren_wide": ("config 5: SirenNet 3 -> 1024 x 8 -> 1 (w0 30) on a synthetic 512^3 volume, coords in [-1,1], "
                   "tensor-core split-precision mode bf16x3 (fp32 parity), Adam lr 1e-4", 17),
    "siren_ankle": ("config 1: SirenNet 4 -> 256 x 5 -> 1 (w0 30) on sample_ankle_dyn_mri.nii.gz, tensor-core "
                    "split-precision mode bf16x3 (fp32 parity), Adam lr 1e-4", 18),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch-log2", type=int, default=None, help="log2 coordinates per step per GPU (default per workload)")
    ap.add_argument("--workload", default="ankle_hash", choices=list(WORKLOADS),
                    help="ankle_hash = BASELINE configs[1] (default, the driver's bench line); the others are configs 1/4/5")
    ap.add_argument("--cpu-batch-log2", type=int, default=15)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-infer", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.batch_log2 is None:
        args.batch_log2 = WORKLOADS[args.workload][1]
    return args


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (subprocess writing to a temp file:
    no Python reader thread competing with the launch loop for the GIL)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None

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
            lines = open(self.path).read().splitlines()[getattr(self, "skip", 0):]
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
def oracle_step_fn(batch_log2: int):
    """One training step of the reference's PyTorch CPU path (oracle port): G4 + 2x64 GELU decoder + Adam."""
    import torch.nn.functional as F
    from oracle import networks, sweep as osweep
    from mri_interpolation_b200 import nifti
    torch.manual_seed(1337)
    params, levels = networks.hashmlp_init(dim_in=4, dim_hidden=64, dim_out=1, n_layers=2, **G4)
    params = {k: v.requires_grad_() for k, v in params.items() if not k.startswith("layers.")}
    opt = torch.optim.Adam(list(params.values()), lr=5e-3)
    vol = torch.from_numpy(nifti.load(SAMPLE).get_fdata(np.float32))
    pixels = osweep.normalise_intensities(vol)
    axes = [osweep.axis_values(s) for s in vol.shape]
    shape = torch.tensor(vol.shape)
    gen = torch.Generator().manual_seed(1337)
    n = 1 << batch_log2

    def step():
        idx = torch.randint(0, pixels.shape[0], (n,), generator=gen)
        rem, cols = idx.clone(), []
        for d in range(3, -1, -1):
            cols.append(axes[d][rem % shape[d]])
            rem = rem // shape[d]
        x = torch.stack(cols[::-1], dim=-1)
        opt.zero_grad()
        loss = F.mse_loss(pixels[idx], networks.hashmlp_forward(x, params, levels, 2, False))
        loss.backward()
        opt.step()
        return float(loss.detach())

    return step, n


def time_oracle(batch_log2: int, steps: int, warmup: int):
    torch.set_num_threads(os.cpu_count() or 1)
    step, n = oracle_step_fn(batch_log2)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return n / dt, dt, n


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    if args.workload != "ankle_hash":
        print(json.dumps({"impl": "reference", "unavailable": "the CPU reference arm is implemented for the default "
                          "workload (ankle_hash, BASELINE configs[1]) only"}))
        return
    steps, warm = max(1, min(args.steps, 5)), max(1, min(args.warmup, 1))
    value, dt, n = time_oracle(args.cpu_batch_log2, steps, warm)
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": "train_coords_per_s", "value": value, "unit": "coords/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "sample_ankle_dyn_mri.nii.gz (bundled) + random-init weights",
        "config": workload_config(args, cpu_sample=n),
        "cpu_baseline": {"value": value, "unit": "coords/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} steps of 2^{args.cpu_batch_log2} coords (oracle port of the reference's "
                                   f"PyTorch CPU path: hash G4 + decoder + MSE + torch.optim.Adam)"},
        "e2e": {"value": value, "unit": "coords/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, cpu_sample=None):
    cfg = {"workload": WORKLOADS[args.workload][0], "workload_name": args.workload,
           "batch_per_gpu": 1 << args.batch_log2, "global_batch": (1 << args.batch_log2) * args.gpus,
           "parallelism": f"dp{args.gpus}" + (" (fused reduce-scatter + Adam + all-gather kernel over NVLink peer memory; "
                                                 "NCCL all-reduce when symmetric memory is unavailable)"
                                                 if args.gpus > 1 else ""), "l2": "inputs larger than L2: every step streams the whole p/g/m/v arena through Adam "
                                                  "(hash: 489 MB/step) plus per-batch activations (SIREN: > 1 GB/layer)"}
    cfg["allocator_priming_steps_before_warmup"] = PRIMING_STEPS
    cfg["batch_order"] = os.environ.get("MRI_BATCH_ORDER", "axis0") + (": i.i.d. uniform voxel indices, each batch arranged with the "
                                                                        "axis-0 index fastest (same sets; the loss is order-invariant)")
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
    out += 0.04 * torch.rand(shape, device=dev, generator=g)
    out = (out - out.min()) / (out.max() - out.min())
    return out.flatten()


def build_workload(args, dev, rank):
    from mri_interpolation_b200 import models, nifti
    from mri_interpolation_b200 import functional as Fn
    torch.manual_seed(1337)
    name = args.workload
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


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    from mri_interpolation_b200 import _lib, distributed, sweep
    from mri_interpolation_b200.datamodules import PrefetchLoader

    rank, local_rank, world = distributed.init_from_env("nccl")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    import torch.distributed as dist

    model, sampler, info = build_workload(args, dev, rank)
    is_hash = args.workload.endswith("hash")
    opt = model.configure_optimizers()
    # bucketed/overlapped all-reduce is available but measured SLOWER at W=2 (1.608 vs 1.561 ms/step): the scatter
    # backward and NCCL's copy kernels both saturate the L2, so the default stays one all-reduce after backward
    if is_hash and world > 1 and os.environ.get("MRI_DP_OVERLAP") == "1":
        opt.enable_overlap(model.encoder, n_groups=4)
    n = 1 << args.batch_log2
    dim = len(info["shape"])
    gen = torch.Generator(device=dev)
    gen.manual_seed(1337 + rank)
    total_steps = args.steps + args.warmup
    # the shuffled index stream is drawn up front (like a DataLoader sampler), as a ring of 32 batches
    ring = 32
    index = torch.randint(0, sampler.total, (ring, n), device=dev, generator=gen)
    # locality-ordered batches (the loaders' default, datamodules.DeviceBatchLoader / functional.locality_sort): the SAME
    # random voxel sets, arranged inside a batch with the axis-0 index fastest.  MRI_BATCH_ORDER=none keeps the drawn order.
    from mri_interpolation_b200 import functional as Fn
    batch_order = os.environ.get("MRI_BATCH_ORDER", "axis0")
    if batch_order != "none":
        index = Fn.locality_sort(index, info["shape"], block=int(batch_order[3:]) if batch_order.startswith("blk") else 1)
    # what drawing + ordering one batch costs on the device (index generation is loader work outside the timed step;
    # DeviceBatchLoader amortises it into one stable sort per epoch)
    torch.cuda.synchronize()
    _t0 = time.perf_counter()
    for _ in range(8):
        _i = torch.randint(0, sampler.total, (n,), device=dev, generator=gen)
        if batch_order != "none":
            _i = Fn.locality_sort(_i, info["shape"], block=1)
    torch.cuda.synchronize()
    sampler_ms = (time.perf_counter() - _t0) / 8 * 1e3

    def step(i):
        x, y = sampler.batch(index[i % ring])
        loss = model.training_step((x, y), i)
        loss.backward()
        opt.step()
        opt.zero_grad()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()  # nvidia-smi needs a few hundred ms to come up: started before the untimed steps, marked below
    for i in range(PRIMING_STEPS):  # untimed: lets torch's caching allocator reach its steady state (no cudaMalloc later)
        step(i)
    for i in range(args.warmup):
        step(i)
    barrier()
    clocks.mark()
    launches0 = _lib.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    inflight = []
    trace = [] if os.environ.get("MRI_BENCH_TRACE") == "1" else None
    for i in range(args.warmup, total_steps):
        loss = step(i)
        if trace is not None:
            te = torch.cuda.Event(enable_timing=True)
            te.record()
            trace.append(te)
        # bound the host's run-ahead to two steps (a saturated launch queue starves NCCL's progress thread and made
        # multi-GPU SIREN steps ~25% slower); the GPU stays fed: the event waited on is two steps old
        ev = torch.cuda.Event()
        ev.record()
        inflight.append(ev)
        if len(inflight) > 2:
            inflight.pop(0).synchronize()
    ev1.record()
    barrier()
    launches = _lib.launch_count - launches0
    ms = ev0.elapsed_time(ev1)
    if trace:
        sys.stderr.write("per-step ms: " + " ".join(f"{a.elapsed_time(b):.1f}" for a, b in zip([ev0] + trace[:-1], trace)) + "\n")
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    # a timed region shorter than nvidia-smi's sampling period holds no clock sample: run the same steps again, untimed,
    # for ~0.5 s on every rank (same count everywhere: the steps are collective) and sample the clocks under that load
    need_probe = torch.tensor([1 if (rank == 0 and clk is not None and clk.get("samples") == 0 and
                                     "nvidia-smi unavailable" not in clk.get("reasons", []) and
                                     os.environ.get("MRI_BENCH_NO_CLOCKS") != "1") else 0], device=dev)
    if world > 1:
        dist.broadcast(need_probe, src=0)
    if int(need_probe.item()):
        probe = ClockSampler(local_rank)
        if rank == 0:
            probe.start()
            time.sleep(0.4)
        barrier()
        probe.mark()
        for i in range(total_steps, total_steps + max(20, int(500.0 / max(ms_step, 1e-3)))):
            step(i)
        barrier()
        if rank == 0:
            clk = probe.stop()
            clk["window"] = "identical untimed steps run right after the timed region (it was shorter than one sample)"
    value = n * world / (ms_step * 1e-3)
    final_loss = float(loss.detach())

    # ---- e2e: host batches through the public API (PrefetchLoader + LightningModule.training_step + FusedAdam)
    # every step: H2D copy of that step's batch from pinned host memory (overlapped with the previous step's
    # compute on a side stream) and a D2H read of that step's loss (async copy, consumed one step later).
    host_ring = []
    for r in range(4):
        xb, yb = sampler.batch(index[r])
        host_ring.append((xb.cpu().pin_memory(), yb.cpu().pin_memory()))
    e2e_steps = max(10, min(args.steps, 100))

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
    for i, (xb, yb) in enumerate(PrefetchLoader(_HostBatches(), dev) if not args.no_e2e else []):
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
    if t0 is None:
        t0 = time.perf_counter() - 1.0
    e2e_dt = torch.tensor([(time.perf_counter() - t0) / e2e_steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_dt, op=dist.ReduceOp.MAX)
    e2e_value = n * world / float(e2e_dt.item())

    # ---- inference: every rank sweeps its slab of the query volume (no communication), max over ranks
    infer = None
    if not args.no_infer:
        if is_hash:
            sweep_shape = SWEEP_SHAPE if args.workload == "ankle_hash" else (256, 256, 256, 4)
        else:
            sweep_shape = (352, 352, 6, 29) if args.workload == "siren_ankle" else (256, 256, 128)
        total_vox = int(np.prod(sweep_shape))
        ns = not is_hash
        for _ in range(2):
            sweep.dense_sweep(model, sweep_shape, norm_siren=ns, rank=rank, world_size=world)
        barrier()
        reps, per_rep = 5, []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = sweep.dense_sweep(model, sweep_shape, norm_siren=ns, rank=rank, world_size=world)
            b.record()
            barrier()
            per_rep.append(a.elapsed_time(b))
        # median of the per-sweep device times: single sweeps occasionally stall for tens of ms on these shared boxes
        sw = torch.tensor([float(np.median(per_rep))], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(sw, op=dist.ReduceOp.MAX)
        sw_ms = float(sw.item())
        t1 = time.perf_counter()
        host = out.cpu()
        d2h_s = time.perf_counter() - t1
        infer = {"metric": "infer_voxels_per_s", "value": total_vox / (sw_ms * 1e-3), "unit": "voxels/s", "n_gpus": world,
                 "workload": f"dense sweep of {sweep_shape} ({total_vox} voxels), contiguous slab per GPU, "
                             + ("fused hash+decoder kernel" if is_hash else "coordinate synthesis + tensor-core SIREN"),
                 "ms": sw_ms, "ms_per_sweep": [round(v, 3) for v in per_rep],
                 "e2e_value_with_d2h": total_vox / (sw_ms * 1e-3 + d2h_s)}
        del host, out

    # ---- isolated kernel timings for the roofline (rank 0, L2 flushed between launches)
    roof, kern = None, None
    if is_hash:
        opt.disable_overlap(model.encoder)
    opt.data_parallel = False  # the isolated-kernel section below runs on rank 0 only: no collectives in it
    if rank == 0:
        hbm_peak, peak_src = peaks()
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
            # the whole-model backward (ONE kernel when the decoder backward is fused with the scatter)
            fused_ms = fused_fwd_ms = None
            if getattr(model, "fuse_backward", False):
                pred = model(x)
                gy = torch.randn_like(pred) * 1e-3
                if pred.grad_fn is not None and type(pred.grad_fn).__name__.startswith("HashDecoderFn"):
                    fused_ms = timed(lambda: torch.autograd.backward(pred, gy, retain_graph=True))
                    fused_fwd_ms = timed(lambda: model(x))  # training forward: gather + decoder, enc and pre2 written
                del pred
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
            traffic = 320.7e6 if top == "hashgrid_bwd" else 235.2e6
            if fused_ms is not None:
                # fused decoder-backward + scatter: coords + enc + dy read, L*2^D*F*4 B reduced into the tables
                fused_bytes = (HASH_BYTES_PER_COORD + 4) * n
                kern["hashdecoder_bwd"] = {"ms": fused_ms, "GBps_algorithmic": fused_bytes / fused_ms / 1e6,
                                           "frac": fused_bytes / fused_ms / 1e6 / hbm_peak}
                kern["hashdecoder_fwd"] = {"ms": fused_fwd_ms, "GBps_algorithmic": (fused_bytes + 4 * n) / fused_fwd_ms / 1e6,
                                           "frac": (fused_bytes + 4 * n) / fused_fwd_ms / 1e6 / hbm_peak}
                if fused_ms >= max(fwd_ms, bwd_ms):
                    top, hash_bytes, traffic = "hashdecoder_bwd", fused_bytes, FUSED_BWD_DRAM_BYTES
            roof = {"kernel": top, "bound": "hbm", "achieved": kern[top]["GBps_algorithmic"], "peak": hbm_peak, "unit": "GB/s",
                    "frac": kern[top]["frac"], "traffic": traffic,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": hash_bytes,
                    "note": "2192 B/coord (4D + L*2^D*F*4 gathered-or-reduced + L*F*4 encoding) x 2^19 coords (+4 B/coord dy for "
                            "the fused decoder-backward+scatter kernel, which also does 8.4 kMAC/coord of decoder math); traffic "
                            "= ncu dram read+write per launch (profiles/r01_ncu_full_hash_adam.csv): the 61 MB of tables stay "
                            "in the 126 MB L2, the kernels are bound by the L1/L2 sector and L2 atomic rates, not HBM"}
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
            tf_peak = 1649.2
            pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
            if os.path.isfile(pk):
                tf_peak = float(json.load(open(pk))["bf16_tflops"])
            issued = 3 * 2.0 * n * h * h
            kern = {"siren_tc_layer_fwd": {"ms": ms_l, "issued_bf16_TFLOPs": issued / ms_l / 1e9, "frac": issued / ms_l / 1e9 / tf_peak},
                    "siren_tc_wgrad": {"ms": ms_w, "issued_bf16_TFLOPs": issued / ms_w / 1e9, "frac": issued / ms_w / 1e9 / tf_peak}}
            roof = {"kernel": "siren_tc_layer_kernel (hidden layer forward, bf16x3 split precision, sine epilogue)",
                    "bound": "tensor", "achieved": kern["siren_tc_layer_fwd"]["issued_bf16_TFLOPs"], "peak": tf_peak,
                    "unit": "TFLOP/s", "frac": kern["siren_tc_layer_fwd"]["frac"], "traffic": None,
                    "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops, burst)",
                    "algorithmic_flops_per_launch": 2.0 * n * h * h,
                    "note": "issued = 3 tcgen05.mma passes x 2 n H^2 (A_lo*B_hi + A_hi*B_lo + A_hi*B_hi); the fp32-equivalent "
                            "(algorithmic) rate is one third of it"}
            if info.get("train_flops_per_coord"):
                kern["whole_step_algorithmic_TFLOPs"] = info["train_flops_per_coord"] * n / (ms_step * 1e-3) / 1e12

    cpu = None
    if rank == 0 and not args.no_cpu_baseline and args.workload == "ankle_hash":
        v, dt, ncpu = time_oracle(args.cpu_batch_log2, 3, 1)
        cpu = {"value": v, "unit": "coords/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"3 steps of 2^{args.cpu_batch_log2} coords of the same training step (oracle port of the "
                         f"reference's PyTorch CPU path), {dt * 1e3:.0f} ms/step"}

    if rank == 0:
        line = {
            "metric": "train_coords_per_s", "value": value, "unit": "coords/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if is_hash else "f32 via bf16x3 split on tcgen05 (fp32 accumulate)",
            "data": info["data"], "config": workload_config(args), "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "coords/s", "h2d_bytes_per_step": n * (dim + 1) * 4, "d2h_bytes_per_step": 4,
                    "steps": e2e_steps},
            "gpu_launches": launches, "final_loss": final_loss, "sampler_ms_per_batch": sampler_ms,
            "roofline": roof, "kernels": kern, "cpu_baseline": cpu, "infer": infer,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
