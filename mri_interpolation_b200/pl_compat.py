"""``pytorch_lightning`` when it is installed, otherwise a minimal stand-in with the surface the
reference's callers use (launcher.py:156-165,179,213; test_script.py:81-91; models.py:20-95):

  pl.LightningModule      nn.Module + .device/.log/.logger/.trainer/.load_from_checkpoint + hooks
  pl.LightningDataModule  prepare_data/setup/*_dataloader
  pl.Trainer(accelerator, devices, max_epochs, accumulate_grad_batches, precision, ...)
      .fit(model, train_dataloaders | datamodule)     one optimiser, automatic optimisation
      .predict(model, dataloaders | datamodule)       -> list of per-batch outputs (inference mode)
      .logger.log_dir / .logger.version               lightning_logs/version_N like Lightning's default
      .save_checkpoint(path)                          {"state_dict": ..., "epoch": ..., "global_step": ...}

The stand-in is orchestration only (L3 in SURVEY.md): no arithmetic happens here.
"""
from __future__ import annotations

import csv
import os
import time
import types
from typing import Any, Dict, Iterable, List, Optional

import torch
from torch import nn

try:  # pragma: no cover - not installed in the build image
    import pytorch_lightning as _real_pl

    HAVE_LIGHTNING = True
except Exception:  # noqa: BLE001
    _real_pl = None
    HAVE_LIGHTNING = False


class _Logger:
    """Tiny CSV logger with Lightning's directory convention."""

    def __init__(self, save_dir: str = ".", name: str = "lightning_logs", version: Optional[int] = None):
        self.save_dir, self.name = save_dir, name
        self._version = version
        self._rows: List[Dict[str, Any]] = []
        self._dir_made = False

    @property
    def version(self) -> int:
        if self._version is None:
            root = os.path.join(self.save_dir, self.name)
            existing = []
            if os.path.isdir(root):
                for d in os.listdir(root):
                    if d.startswith("version_") and d[8:].isdigit():
                        existing.append(int(d[8:]))
            self._version = max(existing) + 1 if existing else 0
        return self._version

    @property
    def log_dir(self) -> str:
        path = os.path.join(self.save_dir, self.name, f"version_{self.version}")
        if not self._dir_made:
            os.makedirs(path, exist_ok=True)
            self._dir_made = True
        return path

    def log_metrics(self, metrics: Dict[str, float], step: int) -> None:
        # device scalars are kept as (cloned) tensors and read back in finalize(): a float() here would drain the GPU's
        # launch queue every 50 steps and at every epoch end
        self._rows.append({"step": step, **{k: (v.detach().clone() if isinstance(v, torch.Tensor) else v) for k, v in metrics.items()}})

    def finalize(self) -> None:
        if not self._rows:
            return
        self._rows = [{k: (float(v) if isinstance(v, torch.Tensor) else v) for k, v in r.items()} for r in self._rows]
        keys: List[str] = []
        for r in self._rows:
            for k in r:
                if k not in keys:
                    keys.append(k)
        with open(os.path.join(self.log_dir, "metrics.csv"), "w", newline="") as f:
            w = csv.DictWriter(f, fieldnames=keys)
            w.writeheader()
            w.writerows(self._rows)


class LightningModule(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        self.__dict__["trainer"] = None
        self.__dict__["_logged"] = {}

    # -- conveniences Lightning offers and the reference relies on
    @property
    def device(self) -> torch.device:
        for p in self.parameters():
            return p.device
        for b in self.buffers():
            return b.device
        return torch.device("cpu")

    @property
    def logger(self):
        return self.trainer.logger if self.trainer is not None else None

    @property
    def current_epoch(self) -> int:
        return self.trainer.current_epoch if self.trainer is not None else 0

    @property
    def global_step(self) -> int:
        return self.trainer.global_step if self.trainer is not None else 0

    def log(self, name: str, value, *args, **kwargs) -> None:
        # keep the device scalar; it is read (one sync) only when metrics are flushed
        if isinstance(value, torch.Tensor):
            value = value.detach()
        self._logged[name] = value
        if self.trainer is not None:
            self.trainer._record(name, value)

    def log_dict(self, d: Dict[str, Any], *args, **kwargs) -> None:
        for k, v in d.items():
            self.log(k, v)

    def save_hyperparameters(self, *args, **kwargs) -> None:
        pass

    def optimizers(self):
        return self.trainer.optimizer if self.trainer is not None else None

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path: str, map_location=None, strict: bool = True, **kwargs):
        ckpt = torch.load(checkpoint_path, map_location=map_location or "cpu", weights_only=False)
        model = cls(**kwargs)
        model.load_state_dict(ckpt["state_dict"] if "state_dict" in ckpt else ckpt, strict=strict)
        return model

    # -- hooks (no-ops)
    def on_fit_start(self): ...
    def on_train_start(self): ...
    def on_train_end(self): ...
    def on_train_epoch_start(self): ...
    def on_train_epoch_end(self): ...
    def on_predict_start(self): ...
    def on_predict_end(self): ...

    def training_step(self, batch, batch_idx):
        raise NotImplementedError

    def predict_step(self, batch, batch_idx):
        return self(batch)

    def configure_optimizers(self):
        raise NotImplementedError


class LightningDataModule:
    def __init__(self, *args, **kwargs):
        pass

    def prepare_data(self): ...
    def setup(self, stage: Optional[str] = None): ...
    def train_dataloader(self): ...
    def val_dataloader(self): ...
    def test_dataloader(self): ...
    def predict_dataloader(self): ...


def _move(batch, device):
    if isinstance(batch, torch.Tensor):
        return batch if batch.device == device else batch.to(device, non_blocking=True)
    if isinstance(batch, (list, tuple)):
        return type(batch)(_move(b, device) for b in batch)
    if isinstance(batch, dict):
        return {k: _move(v, device) for k, v in batch.items()}
    return batch


def training_step_and_backward(model, batch, batch_idx):
    """What Lightning's loop does between two optimiser steps: ``loss = model.training_step(batch)``, ``loss.backward()``.
    A module that offers ``fused_training_step`` (HashMLP under the MSE loss: forward, loss and backward in ONE kernel)
    is asked first; it returns None when it has no fused path for this batch.  Returns the (detached) loss."""
    fused = getattr(model, "fused_training_step", None)
    if fused is not None:
        loss = fused(batch, batch_idx)
        if loss is not None:
            return loss
    loss = model.training_step(batch, batch_idx)
    if isinstance(loss, dict):
        loss = loss["loss"]
    loss.backward()
    return loss.detach()


class Trainer:
    def __init__(self, accelerator: str = "auto", devices: Any = "auto", max_epochs: Optional[int] = None,
                 accumulate_grad_batches: Any = None, precision: Any = 32, logger: Any = True,
                 default_root_dir: Optional[str] = None, enable_progress_bar: bool = False, max_steps: int = -1,
                 callbacks: Optional[list] = None, gpus: Any = None, enable_checkpointing: bool = True,
                 cuda_graph: Optional[bool] = None, **kwargs):
        if str(precision) not in ("32", "32-true"):
            raise NotImplementedError("the B200 backend trains in fp32 (launcher.py:162 precision=32)")
        self.accelerator = accelerator
        self.max_epochs = 1000 if max_epochs is None else max_epochs
        self.max_steps = max_steps
        self.accumulate_grad_batches = accumulate_grad_batches
        self.enable_progress_bar = enable_progress_bar
        self.enable_checkpointing = enable_checkpointing
        # replay the whole training step as one CUDA graph (graph.GraphedTrainStep): for the reference's small batches
        # (4 096 / 10 000 coordinates) the step is host-bound otherwise.  None -> the MRI_CUDA_GRAPH environment switch
        self.cuda_graph = (os.environ.get("MRI_CUDA_GRAPH", "0") == "1") if cuda_graph is None else bool(cuda_graph)
        self.graphed_steps = 0
        root = default_root_dir or os.getcwd()
        if logger is True:
            self.logger = _Logger(save_dir=root)
        elif logger in (False, None):
            self.logger = None
        else:
            self.logger = logger
        self.current_epoch = 0
        self.global_step = 0
        self.optimizer = None
        self.model = None
        self.callback_metrics: Dict[str, Any] = {}
        self._pending: Dict[str, Any] = {}
        self.fit_seconds = 0.0

    # ------------------------------------------------------------------ helpers
    def _device(self) -> torch.device:
        acc = self.accelerator
        if acc in ("gpu", "cuda") or (acc == "auto" and torch.cuda.is_available()):
            if not torch.cuda.is_available():
                raise RuntimeError("accelerator='gpu' requested but CUDA is not available")
            return torch.device("cuda", torch.cuda.current_device())
        return torch.device("cpu")

    def _record(self, name, value) -> None:
        self._pending[name] = value
        self.callback_metrics[name] = value

    def _flush(self) -> None:
        if self.logger is None or not self._pending:
            self._pending = {}
            return
        row = dict(self._pending)  # tensors stay on the device until the logger is finalised (no sync in the loop)
        self.logger.log_metrics(row, self.global_step)
        self._pending = {}

    def _accum_factor(self, epoch: int) -> int:
        a = self.accumulate_grad_batches
        if a is None:
            return 1
        if isinstance(a, int):
            return max(a, 1)
        factor = 1
        for start in sorted(a):  # {epoch: factor} schedule
            if epoch >= int(start):
                factor = int(a[start])
        return max(factor, 1)

    @staticmethod
    def _first_optimizer(cfg):
        if isinstance(cfg, dict):
            cfg = cfg["optimizer"]
        if isinstance(cfg, (list, tuple)):
            cfg = cfg[0]
            if isinstance(cfg, (list, tuple)):
                cfg = cfg[0]
        return cfg

    # ---------------------------------------------------------------------- fit
    def fit(self, model, train_dataloaders=None, val_dataloaders=None, datamodule=None, ckpt_path=None):
        if train_dataloaders is None and datamodule is not None:
            datamodule.prepare_data()
            datamodule.setup()
            train_dataloaders = datamodule.train_dataloader()
        if train_dataloaders is None:
            raise ValueError("fit() needs train_dataloaders or a datamodule")
        device = self._device()
        model.__dict__["trainer"] = self
        self.model = model
        model.to(device)
        model.train()
        if ckpt_path:
            model.load_state_dict(torch.load(ckpt_path, map_location=device, weights_only=False)["state_dict"])
        self.optimizer = self._first_optimizer(model.configure_optimizers())
        if (os.environ.get("MRI_DP_OVERLAP") == "1" and self.accumulate_grad_batches is None
                and hasattr(self.optimizer, "enable_overlap") and hasattr(getattr(model, "encoder", None), "tables")):
            self.optimizer.enable_overlap(model.encoder)  # opt-in: bucketed all-reduce overlapped with the backward
        model.on_fit_start()
        model.on_train_start()
        t0 = time.time()
        stop = False
        graphed = None
        distributed = torch.distributed.is_available() and torch.distributed.is_initialized() and \
            torch.distributed.get_world_size() > 1
        can_graph = (self.cuda_graph and device.type == "cuda" and not distributed
                     and hasattr(self.optimizer, "use_device_step"))
        for epoch in range(self.max_epochs):
            self.current_epoch = epoch
            accum = self._accum_factor(epoch)
            model.on_train_epoch_start()
            self.optimizer.zero_grad(set_to_none=False)
            pending = 0
            for batch_idx, batch in enumerate(train_dataloaders):
                batch = _move(batch, device)
                if can_graph and accum == 1:
                    if graphed is None and isinstance(batch, (tuple, list)) and all(isinstance(t, torch.Tensor) for t in batch):
                        from .graph import GraphedTrainStep
                        graphed = GraphedTrainStep(model, self.optimizer, batch)
                    if graphed is not None and graphed.matches(batch):
                        graphed(batch)
                        for name, value in graphed.logged.items():
                            self._record(name, value)
                        self.graphed_steps += 1
                        self.global_step += 1
                        if self.global_step % 50 == 0:
                            self._flush()
                        if 0 < self.max_steps <= self.global_step:
                            stop = True
                            break
                        continue
                if accum == 1:
                    training_step_and_backward(model, batch, batch_idx)
                else:
                    loss = model.training_step(batch, batch_idx)
                    if isinstance(loss, dict):
                        loss = loss["loss"]
                    (loss / accum).backward()
                pending += 1
                if pending == accum:
                    self.optimizer.step()
                    self.optimizer.zero_grad(set_to_none=False)
                    pending = 0
                    self.global_step += 1
                    if self.global_step % 50 == 0:
                        self._flush()
                    if 0 < self.max_steps <= self.global_step:
                        stop = True
                        break
            if pending and not stop:
                self.optimizer.step()
                self.optimizer.zero_grad(set_to_none=False)
                self.global_step += 1
            self._flush()
            model.on_train_epoch_end()
            if stop:
                break
        if device.type == "cuda":
            torch.cuda.synchronize(device)
        self.fit_seconds = time.time() - t0
        model.on_train_end()
        if self.logger is not None:
            self.logger.finalize()
            if self.enable_checkpointing:
                ckpt_dir = os.path.join(self.logger.log_dir, "checkpoints")
                os.makedirs(ckpt_dir, exist_ok=True)
                self.save_checkpoint(os.path.join(ckpt_dir, f"epoch={self.current_epoch}-step={self.global_step}.ckpt"))
        return None

    # ------------------------------------------------------------------ predict
    def predict(self, model=None, dataloaders=None, datamodule=None, return_predictions: bool = True, ckpt_path=None):
        model = model or self.model
        if dataloaders is None and datamodule is not None:
            dataloaders = datamodule.predict_dataloader()
        if dataloaders is None:
            raise ValueError("predict() needs dataloaders or a datamodule")
        device = self._device()
        model.__dict__["trainer"] = self
        model.to(device)
        was_training = model.training
        model.eval()
        outs = []
        model.on_predict_start()
        with torch.inference_mode():
            for batch_idx, batch in enumerate(dataloaders):
                outs.append(model.predict_step(_move(batch, device), batch_idx))
        model.on_predict_end()
        model.train(was_training)
        return outs

    def save_checkpoint(self, path: str) -> None:
        torch.save({"state_dict": {k: v.detach().cpu() for k, v in self.model.state_dict().items()},
                    "epoch": self.current_epoch, "global_step": self.global_step}, path)


if HAVE_LIGHTNING:  # pragma: no cover
    pl = _real_pl
else:
    pl = types.ModuleType("pytorch_lightning_compat")
    pl.LightningModule = LightningModule
    pl.LightningDataModule = LightningDataModule
    pl.Trainer = Trainer
    pl.__doc__ = __doc__
