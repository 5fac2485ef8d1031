"""Dense-grid interpolation CLI.

The reference's ``interp.py`` (interp.py:35-52) is the *baseline* the networks are compared with: drop every other
frame and re-interpolate linearly in time with ITK in a per-voxel Python loop.  The network-side dense sweep lives
in the reference's ``launcher.py:191-222``.  This script does both on the B200 backend:

  python interp.py sweep --checkpoint CKPT --model_class HashMLP --shape 352 352 6 29 [--out interp.nii.gz]
  python interp.py linear --image data/sample_ankle_dyn_mri.nii.gz [--out itk_interpolated.nii.gz]

``sweep`` queries a fitted model on the dense grid (volume sharded in slabs over the ranks under torchrun, no
communication); ``linear`` is the drop-odd-frames / linear re-interpolation baseline with its PSNR against the truth.
"""
import argparse
import os

import numpy as np
import torch

from mri_interpolation_b200 import config as base
from mri_interpolation_b200 import distributed, metrics, models, nifti, sweep


def linear_time_baseline(data) -> torch.Tensor:
    """interp.py:35-50: keep frames ::2, linear interpolation at continuous index t/2 along the last axis
    (csrc/metrics.cu::linear_time_kernel; returns a CUDA tensor)."""
    return metrics.linear_time_baseline(data)


def main(argv=None):
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="cmd", required=True)
    s = sub.add_parser("sweep")
    s.add_argument("--checkpoint", required=True)
    s.add_argument("--model_class", default="HashMLP")
    s.add_argument("--shape", type=int, nargs="+", required=True)
    s.add_argument("--norm_siren", action="store_true")
    s.add_argument("--out", default="interpolation.nii.gz")
    l = sub.add_parser("linear")
    l.add_argument("--image", default=base.BaseConfig().image_path)
    l.add_argument("--slice", type=int, default=3, help="z slice for 4-D volumes (the reference uses data[:, :, 3, :])")
    l.add_argument("--out", default="itk_interpolated.nii.gz")
    args = ap.parse_args(argv)

    if args.cmd == "linear":
        data = nifti.load(args.image).get_fdata(np.float32)
        data = data / data.max()
        if data.ndim == 4:
            data = data[:, :, args.slice, :]
        truth = torch.from_numpy(np.ascontiguousarray(data)).cuda()
        interpolated = linear_time_baseline(truth)
        nifti.save(interpolated.cpu().numpy().astype(np.float32), args.out)
        odd = slice(1, None, 2)
        print(f"linear-in-time baseline: PSNR on the re-interpolated (odd) frames "
              f"{metrics.peak_signal_noise_ratio(truth[..., odd], interpolated[..., odd]):.2f} dB, "
              f"SSIM {metrics.structural_similarity(truth, interpolated):.4f} -> {args.out}")
        return interpolated

    rank, local_rank, world = distributed.init_from_env()
    cfg = base.HashConfig()
    kwargs = dict(dim_in=len(args.shape), dim_hidden=cfg.dim_hidden, dim_out=cfg.dim_out, n_layers=cfg.n_layers,
                  n_levels=cfg.n_levels, n_features_per_level=cfg.n_features_per_level,
                  log2_hashmap_size=cfg.log2_hashmap_size, base_resolution=cfg.base_resolution,
                  finest_resolution=cfg.finest_resolution, lr=cfg.lr)
    model = getattr(models, args.model_class).load_from_checkpoint(args.checkpoint, strict=False, **kwargs)
    model = model.to(torch.device("cuda", local_rank))
    local = sweep.dense_sweep(model, args.shape, norm_siren=args.norm_siren, rank=rank, world_size=world)
    full = sweep.gather_slabs(local, args.shape)
    if rank == 0:
        nifti.save(np.asarray(full, dtype=np.float32), args.out)
        print(f"wrote {args.out} {tuple(args.shape)}")
    return full


if __name__ == "__main__":
    main()
