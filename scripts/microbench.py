#!/usr/bin/env python
"""Per-kernel timings on one B200 (CUDA events, L2 flushed between launches). Writes gpurun_out/microbench.json."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mri_interpolation_b200 import encoding, models, sweep, tc  # noqa: E402
from mri_interpolation_b200 import functional as Fn  # noqa: E402
from mri_interpolation_b200._lib import ACT_GELU, ACT_IDENTITY, ACT_SINE  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
PEAK_HBM = 6547.2
PEAK_BF16 = 1649.2
if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    PEAK_HBM, PEAK_BF16 = pk["hbm_gbs"], pk["bf16_tflops"]


def timed(fn, reps=10, warm=3, do_flush=True):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        if do_flush:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.mean(ts)), float(np.min(ts))


res = {}
G4 = dict(n_levels=16, n_features_per_level=2, log2_hashmap_size=19, base_resolution=16, finest_resolution=2489)
which = sys.argv[1:] or ["hash", "decoder", "sweep", "siren"]

if "hash" in which or "decoder" in which or "sweep" in which:
    torch.manual_seed(1337)
    model = models.HashMLP(dim_in=4, dim_hidden=64, dim_out=1, n_layers=2, lr=5e-3, batch_norm=False, **G4).to(dev)
    with torch.no_grad():
        for lv in model.encoder.levels:
            lv.embedding.weight.uniform_(-0.5, 0.5)
    enc = model.encoder

if "hash" in which:
    for lg in (18, 19, 20, 22):
        n = 1 << lg
        x = torch.rand(n, 4, device=dev)
        go = torch.randn(n, 32, device=dev)
        with torch.no_grad():
            f_ms, _ = timed(lambda: enc(x))
        out = enc(x)
        b_ms, _ = timed(lambda: torch.autograd.backward(out, go, retain_graph=True))
        bytes_ = 2192 * n
        res[f"hashgrid_fwd_random_2^{lg}"] = {"ms": f_ms, "GBps": bytes_ / f_ms / 1e6, "frac_hbm": bytes_ / f_ms / 1e6 / PEAK_HBM}
        res[f"hashgrid_bwd_random_2^{lg}"] = {"ms": b_ms, "GBps": bytes_ / b_ms / 1e6, "frac_hbm": bytes_ / b_ms / 1e6 / PEAK_HBM}
        del out, x, go
    # coherent (sweep-ordered) coordinates
    shape = (352, 352, 6, 29)
    axes = [torch.linspace(0, 1, s) for s in shape]
    n = 1 << 22
    xc = Fn.grid_coords(axes, 5_000_000, n, dev)
    with torch.no_grad():
        f_ms, _ = timed(lambda: enc(xc))
    res["hashgrid_fwd_coherent_2^22"] = {"ms": f_ms, "GBps": 2192 * n / f_ms / 1e6, "frac_hbm": 2192 * n / f_ms / 1e6 / PEAK_HBM}

if "decoder" in which:
    n = 1 << 19
    z = torch.randn(n, 32, device=dev, requires_grad=True)
    l1, l2 = model.decoder[0][0], model.decoder[1][0]
    with torch.no_grad():
        f_ms, _ = timed(lambda: Fn.Decoder2Fn.apply(z, l1.weight, l1.bias, l2.weight, l2.bias, ACT_GELU, ACT_GELU))
    y = Fn.Decoder2Fn.apply(z, l1.weight, l1.bias, l2.weight, l2.bias, ACT_GELU, ACT_GELU)
    gy = torch.randn_like(y)
    b_ms, _ = timed(lambda: torch.autograd.backward(y, gy, retain_graph=True))
    res["decoder2_fwd_2^19"] = {"ms": f_ms, "coords_per_s": n / f_ms * 1e3}
    res["decoder2_bwd_2^19"] = {"ms": b_ms, "coords_per_s": n / b_ms * 1e3}

if "sweep" in which:
    shape = (352, 352, 6, 29)
    tot = int(np.prod(shape))
    f_ms, _ = timed(lambda: sweep.dense_sweep(model, shape), reps=3, warm=1)
    u_ms, _ = timed(lambda: sweep.dense_sweep(model, shape, batch_size=1 << 21, fused=False), reps=3, warm=1)
    res["sweep_fused_21.5M"] = {"ms": f_ms, "voxels_per_s": tot / f_ms * 1e3}
    res["sweep_unfused_21.5M"] = {"ms": u_ms, "voxels_per_s": tot / u_ms * 1e3}

if "siren" in which:
    for (n, h) in ((1 << 17, 1024), (1 << 18, 1024), (1 << 18, 256)):
        a = torch.rand(n, h, device=dev) * 2 - 1
        w = (torch.rand(h, h, device=dev) * 2 - 1) * (6.0 / h) ** 0.5 / 30
        b = torch.zeros(h, device=dev)
        a_hi, a_lo = tc.split(a)
        w_hi, w_lo = tc.split(w)
        flops = 2.0 * n * h * h
        for passes in (3, 1):
            ms, mn = timed(lambda: tc.layer(a_hi, a_lo if passes == 3 else None, w_hi, w_lo if passes == 3 else None, b, ACT_SINE,
                                            30.0, passes=passes, want_planes=True, want_aux=True), reps=5)
            res[f"siren_tc_fwd_n{n}_h{h}_p{passes}"] = {"ms": ms, "algorithmic_TFLOPs": flops / ms / 1e9,
                                                        "issued_bf16_TFLOPs": passes * flops / ms / 1e9,
                                                        "tensor_frac_of_measured_bf16_peak": passes * flops / ms / 1e9 / PEAK_BF16}
            ms, mn = timed(lambda: tc.layer(a_hi, a_lo if passes == 3 else None, w_hi, w_lo if passes == 3 else None, b, ACT_IDENTITY,
                                            1.0, passes=passes, want_planes=True), reps=5)
            res[f"siren_tc_gemm_only_n{n}_h{h}_p{passes}"] = {"ms": ms, "issued_bf16_TFLOPs": passes * flops / ms / 1e9,
                                                              "tensor_frac_of_measured_bf16_peak": passes * flops / ms / 1e9 / PEAK_BF16}
            gw = torch.zeros(h, h, device=dev)
            gb = torch.zeros(h, device=dev)
            ms, mn = timed(lambda: tc.wgrad(a_hi, a_lo if passes == 3 else None, a_hi, a_lo if passes == 3 else None, gw, None,
                                            passes=passes), reps=5)
            res[f"siren_tc_wgrad_n{n}_h{h}_p{passes}"] = {"ms": ms, "issued_bf16_TFLOPs": passes * flops / ms / 1e9,
                                                          "tensor_frac_of_measured_bf16_peak": passes * flops / ms / 1e9 / PEAK_BF16}
        # cuBLAS reference points (library GEMMs, for context only)
        ab, wb = a.to(torch.bfloat16), w.to(torch.bfloat16)
        ms, _ = timed(lambda: torch.matmul(ab, wb.t()), reps=5)
        res[f"cublas_bf16_n{n}_h{h}"] = {"ms": ms, "TFLOPs": flops / ms / 1e9}
        ms, _ = timed(lambda: torch.matmul(a, w.t()), reps=5)
        res[f"cublas_fp32_n{n}_h{h}"] = {"ms": ms, "TFLOPs": flops / ms / 1e9}
        del a, w, a_hi, a_lo, ab, wb
    # whole network, config 5: SirenNet 3 -> 1024 x 8 -> 1
    torch.manual_seed(1337)
    net = models.SirenNet(dim_in=3, dim_hidden=1024, n_layers=8).to(dev)
    opt = net.configure_optimizers()
    n = 1 << 17
    x = torch.rand(n, 3, device=dev) * 2 - 1
    y = torch.rand(n, 1, device=dev)
    for mode in ("bf16x3", "bf16", "fp32"):
        net.precision = mode
        reps = 2 if mode == "fp32" else 5
        with torch.no_grad():
            f_ms, _ = timed(lambda: net(x), reps=reps, warm=1)

        def train():
            loss = net.training_step((x, y), 0)
            loss.backward()
            opt.step()
            opt.zero_grad()
        t_ms, _ = timed(train, reps=reps, warm=1)
        res[f"siren_8x1024_{mode}_n{n}"] = {"fwd_ms": f_ms, "infer_voxels_per_s": n / f_ms * 1e3, "train_ms": t_ms,
                                            "train_coords_per_s": n / t_ms * 1e3,
                                            "fwd_algorithmic_TFLOPs": 14.688e6 * n / f_ms / 1e9,
                                            "train_algorithmic_TFLOPs": 44.06e6 * n / t_ms / 1e9}

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = os.path.join(ROOT, "gpurun_out", "microbench.json")
prev = {}
if os.path.isfile(out):
    try:
        prev = json.load(open(out))
    except Exception:  # noqa: BLE001
        prev = {}
prev.update(res)
json.dump(prev, open(out, "w"), indent=1)
for k, v in res.items():
    print(k, json.dumps(v))

if "levels" in which:
    # per-level cost of the scatter backward and the gather forward (G4, 2^19 random coords)
    from mri_interpolation_b200 import _lib
    torch.manual_seed(1337)
    model = models.HashMLP(dim_in=4, dim_hidden=64, dim_out=1, n_layers=2, lr=5e-3, batch_norm=False, **G4).to(dev)
    enc = model.encoder
    n = 1 << 19
    x = torch.rand(n, 4, device=dev)
    go = torch.randn(n, 32, device=dev)
    tables = enc.tables()
    grads = [torch.zeros_like(t) for t in tables]
    enc._bwd_layout.refresh(grads, enc._resolutions, enc._rows)
    per = {}
    for lv in range(16):
        ms, _ = timed(lambda: _lib.call("mri_hashgrid_backward_levels", x.data_ptr(), n, 4, go.data_ptr(), enc._bwd_layout.base,
                                        enc._bwd_layout.levels, 16, 2, lv, 1, _lib.stream()), reps=5)
        per[f"bwd_level_{lv}_rows_{enc._rows[lv]}"] = round(ms * 1e3, 1)
    res["hashgrid_bwd_per_level_us"] = per
    json.dump({**prev, **res}, open(out, "w"), indent=1)
    print(json.dumps(per, indent=0))

if "sirenlayer" in which:
    # one hidden layer of the 8x1024 SIREN at n = 2^17: forward (sine epilogue), dgrad, wgrad - the ncu target
    n, h = 1 << 17, 1024
    a = torch.rand(n, h, device=dev) * 2 - 1
    w = (torch.rand(h, h, device=dev) * 2 - 1) * (6.0 / h) ** 0.5 / 30
    b = torch.zeros(h, device=dev)
    a_hi, a_lo = tc.split(a)
    w_hi, w_lo = tc.split(w)
    gw = torch.zeros(h, h, device=dev)
    for _ in range(4):
        oh, ol, _, aux = tc.layer(a_hi, a_lo, w_hi, w_lo, b, ACT_SINE, 30.0, passes=3, want_planes=True, want_aux=True)
        tc.wgrad(a_hi, a_lo, oh, ol, gw, None, passes=3)
        tc.dgrad(a_hi, a_lo, w_hi, w_lo, passes=3, mul=aux, want_planes=True)
    torch.cuda.synchronize()
    print("sirenlayer done")
