"""CUDA-graph capture of one whole training step (single GPU).

The reference trains with batches of 4 096 / 10 000 coordinates (config/base.py:23,63).  At that size the five kernels
of a step finish in tens of microseconds and the step is bound by Python, autograd and launch overhead (≈0.5 ms on the
host for 2^13 coordinates).  ``GraphedTrainStep`` records

    loss = model.training_step(batch, 0); loss.backward(); optimizer.step(); optimizer.zero_grad()

once into a ``torch.cuda.CUDAGraph`` over static input buffers and replays it per batch: no Python in the loop, one
``cudaGraphLaunch`` per step.  Everything a replay needs lives in device memory - in particular Adam's step counter
(``FusedAdam.use_device_step`` / ``mri_adam_step_captured``), which a host-side argument would freeze.

Semantics are the eager loop's: the one eager warm-up step that CUDA needs before a capture (lazy kernel attributes,
allocator pools) is rolled back from a snapshot of parameters, gradients, moments, module buffers and the counter.
"""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import torch

from ._lib import MriB200Error


def _key(batch: Sequence[torch.Tensor]) -> Tuple:
    return tuple((tuple(t.shape), t.dtype, t.device) for t in batch)


class GraphedTrainStep:
    """Callable ``batch -> loss`` replaying a captured training step.  The returned loss is a static device tensor that
    the next call overwrites; read it (``float(loss)``) before then if it is needed."""

    def __init__(self, model, optimizer, example_batch: Sequence[torch.Tensor]):
        if not all(isinstance(t, torch.Tensor) and t.is_cuda for t in example_batch):
            raise MriB200Error("GraphedTrainStep: the batch must be a tuple of CUDA tensors")
        if not hasattr(optimizer, "use_device_step"):
            raise MriB200Error("GraphedTrainStep needs the arena optimiser (optim.FusedAdam)")
        self.model, self.optimizer = model, optimizer
        self.static = tuple(t.detach().clone() for t in example_batch)
        self.key = _key(self.static)
        self.logged: Dict[str, torch.Tensor] = {}
        optimizer.use_device_step()

        # eager warm-up on a side stream, then roll every mutation back
        snap = optimizer.snapshot()
        buffers = {name: b.detach().clone() for name, b in model.named_buffers()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._step()
        torch.cuda.current_stream().wait_stream(side)
        optimizer.restore(snap)
        with torch.no_grad():
            for name, b in model.named_buffers():
                b.copy_(buffers[name])

        host_count, host_clean = optimizer.step_count, optimizer._grads_clean
        before = dict(getattr(model, "_logged", {}))
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._step()
        # capturing records the launches without running them: nothing happened on the device yet
        optimizer.step_count, optimizer._grads_clean = host_count, host_clean
        logged = getattr(model, "_logged", {})
        self.logged = {k: v for k, v in logged.items() if isinstance(v, torch.Tensor) and before.get(k) is not v}
        self.replays = 0

    def _step(self) -> torch.Tensor:
        loss = self.model.training_step(self.static, 0)
        if isinstance(loss, dict):
            loss = loss["loss"]
        loss.backward()
        self.optimizer.step()
        self.optimizer.zero_grad(set_to_none=False)
        return loss.detach()

    def matches(self, batch) -> bool:
        return (isinstance(batch, (tuple, list)) and len(batch) == len(self.static)
                and all(isinstance(t, torch.Tensor) for t in batch) and _key(batch) == self.key)

    def __call__(self, batch: Sequence[torch.Tensor]) -> torch.Tensor:
        for dst, src in zip(self.static, batch):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        self.optimizer.note_replayed_step()
        self.replays += 1
        return self.loss
