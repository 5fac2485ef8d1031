"""Launcher for trainings using datamodules and models - the reference's ``launcher.py`` on the B200 backend.

Same flow and flags as the reference (launcher.py:34-224): config -> model -> datamodule -> ``pl.Trainer.fit`` ->
``trainer.predict`` -> pred.nii.gz -> dense-grid interpolations -> config.txt, plus what the reference left
commented out or broken: ``--model_class`` is honoured (launcher.py:82-88), ``interp_shapes`` is iterated
(launcher.py:196 reads a non-existent ``interp_shape``), file paths use ``os.path.join`` and a scores.txt
(MSE / PSNR / SSIM, legacy_code/hash_experimentation.py:445-459) is written.  Under ``torchrun`` every rank trains
data-parallel and sweeps its own slab of each interpolation grid.
"""
import argparse
import json
import os
import time

import numpy as np
import torch

from mri_interpolation_b200 import config as base
from mri_interpolation_b200 import distributed, metrics, models, nifti, sweep
from mri_interpolation_b200.pl_compat import pl

torch.manual_seed(1337)


def parse_args(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--batch_size", help="batch size", type=int, required=False)
    parser.add_argument("--epochs", help="Number of epochs", type=int, required=False)
    parser.add_argument("--accumulate_grad_batches", help="number of batches accumulated per gradient descent step",
                        type=int, required=False)
    parser.add_argument("--n_sample", help="number of points for psf in x, y, z", type=int, required=False)
    parser.add_argument("--model_class", help="Modele class selection", type=str, required=False)
    parser.add_argument("--enco_config_path", help="path of the hash encoding json (config/hash_config.json layout)",
                        type=str, required=False)
    parser.add_argument("--image_path", type=str, required=False)
    parser.add_argument("--max_steps", type=int, default=-1)
    parser.add_argument("--output_dir", type=str, default=None)
    return parser.parse_args(argv)


def build_config(args):
    cfg = base.HashConfig()
    enco_path = args.enco_config_path or os.path.join(os.path.dirname(os.path.abspath(__file__)), "config", "hash_config.json")
    if os.path.isfile(enco_path):
        with open(enco_path) as f:
            cfg.enco_config = json.load(f)
        if args.enco_config_path:  # explicit json: use its grid geometry through the python API
            enc = cfg.enco_config["encoding"]
            cfg.n_levels, cfg.n_features_per_level = enc["n_levels"], enc["n_features_per_level"]
            cfg.log2_hashmap_size, cfg.base_resolution = enc["log2_hashmap_size"], enc["base_resolution"]
            cfg.finest_resolution = round(enc["base_resolution"] * enc["per_level_scale"] ** (enc["n_levels"] - 1))
    for key, value in vars(args).items():
        if value is not None and key not in ("model_class", "enco_config_path", "max_steps", "output_dir"):
            setattr(cfg, key, value)
    if args.image_path:
        cfg.image_shape = nifti.load(args.image_path).shape
        cfg.dim_in = len(cfg.image_shape)
    if args.model_class is not None:
        if not hasattr(models, args.model_class):
            raise SystemExit("model class not recognized, exiting")
        cfg.model_cls = getattr(models, args.model_class)
    return cfg


def build_model(cfg):
    kwargs = dict(dim_in=cfg.dim_in, dim_hidden=cfg.dim_hidden, dim_out=cfg.dim_out, n_layers=cfg.n_layers,
                  encoder_type=cfg.encoder_type, n_levels=cfg.n_levels, n_features_per_level=cfg.n_features_per_level,
                  log2_hashmap_size=cfg.log2_hashmap_size, base_resolution=cfg.base_resolution,
                  finest_resolution=cfg.finest_resolution, per_level_scale=cfg.per_level_scale,
                  interpolation=cfg.interpolation, w0=cfg.w0, w0_initial=cfg.w0_initial, use_bias=cfg.use_bias,
                  final_activation=cfg.final_activation, lr=cfg.lr)
    if cfg.checkpoint_path:
        return cfg.model_cls.load_from_checkpoint(cfg.checkpoint_path, **kwargs)
    return cfg.model_cls(**kwargs)


def main(argv=None):
    args = parse_args(argv)
    rank, local_rank, world = distributed.init_from_env()
    cfg = build_config(args)
    model = build_model(cfg)

    datamodule = cfg.datamodule(config=cfg)
    datamodule.prepare_data()
    datamodule.setup()
    train_loader = datamodule.train_dataloader()

    trainer = pl.Trainer(
        accelerator="gpu" if torch.cuda.is_available() else "cpu",
        max_epochs=cfg.epochs,
        accumulate_grad_batches=cfg.accumulate_grad_batches if cfg.accumulate_grad_batches else None,
        precision=32,
        max_steps=args.max_steps,
        default_root_dir=args.output_dir,
        logger=(rank == 0),
    )
    training_start = time.time()
    trainer.fit(model, train_loader)
    training_stop = time.time()

    filepath = (model.logger.log_dir if rank == 0 else "") + os.sep
    if rank == 0:
        cfg.log = str(model.logger.version)

    # prediction on the training grid = the dense sweep at the image's own shape
    local = sweep.dense_sweep(model, cfg.image_shape, batch_size=max(cfg.batch_size, 1 << 20), rank=rank, world_size=world)
    im = sweep.gather_slabs(local, cfg.image_shape)
    if rank == 0:
        im = np.asarray(im, dtype=np.float32)
        nifti.save(im, filepath + "pred.nii.gz")
        truth = datamodule.dataset.pixels.reshape(cfg.image_shape).numpy()
        n_params = sum(p.numel() for p in model.parameters() if p.requires_grad)
        metrics.write_scores(filepath + "scores.txt", truth, im,
                             {"training time": f"{training_stop - training_start} seconds",
                              "Number of trainable parameters": n_params,
                              "Max memory allocated": torch.cuda.max_memory_allocated() if torch.cuda.is_available() else 0})

    for shape in cfg.interp_shapes:
        t0 = time.time()
        local = sweep.dense_sweep(model, shape, batch_size=max(cfg.batch_size, 1 << 20), rank=rank, world_size=world)
        interp_im = sweep.gather_slabs(local, shape)
        t_sweep = time.time() - t0
        if rank == 0:
            nifti.save(np.asarray(interp_im, dtype=np.float32), filepath + f"interpolation{tuple(shape)}.nii.gz")
            print(f"interpolation {tuple(shape)}: {int(np.prod(shape)) / t_sweep / 1e6:.1f} Mvoxel/s incl. host gather, "
                  f"{time.time() - t0 - t_sweep:.1f} s to gzip and write the NIfTI")

    if rank == 0:
        cfg.export_to_txt(file_path=filepath)
        print(f"outputs in {filepath}")
    return model, trainer


if __name__ == "__main__":
    main()
