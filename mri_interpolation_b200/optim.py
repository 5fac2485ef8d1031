"""Flat parameter arena + fused dense Adam (K6) with the data-parallel gradient all-reduce.

Reference: ``BaseMLP.configure_optimizers`` returns ``torch.optim.Adam(self.parameters(), lr)``
(models.py:68-70) - dense Adam, defaults beta=(0.9, 0.999), eps=1e-8, no weight decay; every
table row's m/v decay every step even with a zero gradient.  ``FusedAdam`` keeps those semantics
but runs ONE kernel over one flat buffer [tables | MLP] (28 B/param, 32 B with the fused
gradient clear), and - when torch.distributed is initialised with world_size > 1 - sums the flat
gradient over the ranks with a single NCCL all-reduce (NVLink 5 / NVSwitch) right before it.
"""
from __future__ import annotations

import ctypes
import os
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist

from . import _lib
from ._lib import MriB200Error
from .distributed import allreduce_sum_

# multimem mode only: 1 = the owner of a slice clears it on every rank with a multicast store (one more arena of NVLink
# ingress per rank and step), 0 = every rank clears its own gradient arena after the closing barrier
_MULTICAST_CLEAR = os.environ.get("MRI_DP_MULTICAST_CLEAR", "0") == "1"

_ALIGN = 4  # floats: every parameter starts 16-byte aligned inside the arena


class FlatArena:
    """Re-homes a list of parameters into one flat fp32 buffer (and their gradients into another),
    keeping every ``nn.Parameter`` object - hence every state_dict key and shape - unchanged."""

    def __init__(self, params: Iterable[torch.nn.Parameter], symmetric_group=None, pad_to: int = _ALIGN):
        """``symmetric_group``: allocate both arenas as torch symmetric memory of that process group so that
        every rank can load/store every other rank's arena over NVLink (peer pointers in ``data_hdl.buffer_ptrs``)."""
        self.params: List[torch.nn.Parameter] = [p for p in params]
        if not self.params:
            raise MriB200Error("FlatArena: no parameters")
        dev = self.params[0].device
        for p in self.params:
            if p.device != dev:
                raise MriB200Error("FlatArena: parameters live on different devices")
            if not p.is_cuda or p.dtype != torch.float32:
                raise MriB200Error(
                    f"FlatArena: parameters must be CUDA float32 (got {p.device}, {p.dtype}); move the model to "
                    f"the GPU before configure_optimizers() - there is no CPU fallback")
        self.offsets, total = [], 0
        for p in self.params:
            self.offsets.append(total)
            total += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        total = (total + pad_to - 1) // pad_to * pad_to
        self.numel = total
        self.had_grads = any(p.grad is not None for p in self.params)  # copied into the arena below: it is not all zeros
        self.data_hdl = self.grad_hdl = None
        if symmetric_group is not None:
            import torch.distributed._symmetric_memory as symm
            self.data = symm.empty(total, dtype=torch.float32, device=dev)
            self.grad = symm.empty(total, dtype=torch.float32, device=dev)
            self.data.zero_()
            self.grad.zero_()
            self.data_hdl = symm.rendezvous(self.data, symmetric_group)
            self.grad_hdl = symm.rendezvous(self.grad, symmetric_group)
        else:
            self.data = torch.zeros(total, device=dev, dtype=torch.float32)
            self.grad = torch.zeros(total, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for p, off in zip(self.params, self.offsets):
                view = self.data[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                old_grad = p.grad
                p.data = view
                gview = self.grad[off:off + p.numel()].view(p.shape)
                if old_grad is not None:
                    gview.copy_(old_grad)
                p.grad = gview

    def intact(self) -> bool:
        """True while every parameter (and its .grad) still aliases the arena (``model.to()`` or
        ``zero_grad(set_to_none=True)`` would break that)."""
        base, gbase = self.data.data_ptr(), self.grad.data_ptr()
        for p, off in zip(self.params, self.offsets):
            if p.data_ptr() != base + 4 * off or p.grad is None or p.grad.data_ptr() != gbase + 4 * off:
                return False
        return True

    def reattach_grads(self) -> None:
        for p, off in zip(self.params, self.offsets):
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * off:
                gview = self.grad[off:off + p.numel()].view(p.shape)
                if p.grad is not None:
                    gview.copy_(p.grad)
                else:
                    gview.zero_()
                p.grad = gview


class FusedAdam(torch.optim.Optimizer):
    """Dense Adam with torch.optim.Adam's semantics, one kernel over a FlatArena.

    ``process_group``: None -> the default group when torch.distributed is initialised.
    ``grad_average``: gradients are summed over ranks and scaled by 1/world_size (each rank computes
    the mean loss of its own shard of the global batch), i.e. the update equals the single-GPU
    update on the concatenated batch up to fp32 summation order.
    """

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 process_group=None, grad_average: bool = True, fuse_zero_grad: bool = True, data_parallel: bool = True,
                 sharded: Optional[bool] = None):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise MriB200Error("FusedAdam handles a single parameter group (the reference uses one)")
        plist = [p for p in self.param_groups[0]["params"] if p.requires_grad]
        self.process_group = process_group
        self.data_parallel = data_parallel  # False: never all-reduce, even when torch.distributed is initialised
        world = self._world()
        if sharded is None:
            sharded = os.environ.get("MRI_DP_SHARDED", "1") == "1"
        self.sharded = False
        if sharded and world > 1:
            # fused reduce-scatter + Adam + all-gather over NVLink peer memory (csrc/optim.cu::adam_sharded_kernel)
            try:
                group = process_group if process_group is not None else dist.group.WORLD
                self.arena = FlatArena(plist, symmetric_group=group, pad_to=4 * world)
                self.sharded = True
            except Exception as e:  # noqa: BLE001 - no P2P / symmetric memory on this box: NCCL all-reduce path
                import warnings
                warnings.warn(f"symmetric-memory arena unavailable ({e}); falling back to the NCCL all-reduce + full Adam step")
        if not self.sharded:
            self.arena = FlatArena(plist)
        if self.sharded:
            self.shard_len = self.arena.numel // world
            self.shard_begin = self.shard_len * dist.get_rank(self.process_group)
            self.exp_avg = torch.zeros(self.shard_len, device=self.arena.data.device, dtype=torch.float32)
            self.exp_avg_sq = torch.zeros_like(self.exp_avg)
            self._peer_grads = (ctypes.c_uint64 * world)(*[int(p) for p in self.arena.grad_hdl.buffer_ptrs])
            self._peer_params = (ctypes.c_uint64 * world)(*[int(p) for p in self.arena.data_hdl.buffer_ptrs])
            # NVSwitch multicast (NVLS) mappings when the fabric offers them: in-switch reduction + multicast store
            # measured (bench ankle_hash): W=2 1.221 ms/step with P2P pointers vs 1.295 with multimem; W=8 1.299 vs 1.285
            mode = os.environ.get("MRI_DP_MULTIMEM", "auto")
            use_mc = mode == "1" or (mode == "auto" and world >= 4)
            self._grad_mc = int(getattr(self.arena.grad_hdl, "multicast_ptr", 0) or 0) if use_mc else 0
            self._param_mc = int(getattr(self.arena.data_hdl, "multicast_ptr", 0) or 0) if use_mc else 0
            if not (self._grad_mc and self._param_mc):
                self._grad_mc = self._param_mc = 0
            # in-kernel synchronisation (csrc/optim.cu, SYNC = true; opt-in with MRI_DP_INKERNEL_SYNC=1): the two
            # symmetric-memory barrier launches around the exchange kernel become flag stores / polls inside it.  Measured
            # at W = 2: 0.804 vs 0.800 ms/step - the barrier launches cost nothing beyond the wait for the slowest rank, which
            # the kernel now does itself (profiles/r02_exchange_ab_w8.txt) - so the default stays the barrier launches.
            self._peer_flags = None
            if os.environ.get("MRI_DP_INKERNEL_SYNC", "0") == "1":
                import torch.distributed._symmetric_memory as symm
                self._flags = symm.empty(32, dtype=torch.int32, device=self.arena.data.device)
                self._flags.zero_()
                self._flags_hdl = symm.rendezvous(self._flags, group)
                self._flags_hdl.barrier(channel=2)  # every rank's flags are zero before anyone's first step
                self._peer_flags = (ctypes.c_uint64 * world)(*[int(p) for p in self._flags_hdl.buffer_ptrs])
        else:
            self.exp_avg = torch.zeros_like(self.arena.data)
            self.exp_avg_sq = torch.zeros_like(self.arena.data)
        self.step_count = 0
        # measurement switch (bench.py): per-step device times of (opening barrier, exchange kernel, closing barrier, clear)
        self.time_exchange = False
        self.exchange_times = []
        self.grad_average = grad_average
        self.fuse_zero_grad = fuse_zero_grad
        # "the gradient arena is all zeros" is tracked with a token, not a guess: functional.grad_write_epoch() moves
        # whenever a backward kernel accumulates into a .grad buffer, post-accumulate hooks catch autograd's own
        # accumulations (BatchNorm parameters etc.), and gradients that existed before the arena was built were copied in
        self._clean_token = None
        self._torch_dirty = False
        for p in self.arena.params:
            p.register_post_accumulate_grad_hook(self._mark_dirty)
        self._grads_clean = not self.arena.had_grads
        self.allreduce_count = 0
        self._overlap = None
        self._step_dev = None   # device-side step counter + {lr/bc1, sqrt(bc2)} scratch: set by use_device_step()
        self._hyper_dev = None

    def _mark_dirty(self, _param=None) -> None:
        self._torch_dirty = True

    @property
    def _grads_clean(self) -> bool:
        from . import functional as Fn
        return self._clean_token is not None and self._clean_token == Fn.grad_write_epoch() and not self._torch_dirty

    @_grads_clean.setter
    def _grads_clean(self, clean: bool) -> None:
        from . import functional as Fn
        self._clean_token = Fn.grad_write_epoch() if clean else None
        if clean:
            self._torch_dirty = False

    # -- distributed
    def _world(self) -> int:
        if self.data_parallel and dist.is_available() and dist.is_initialized():
            return dist.get_world_size(self.process_group)
        return 1

    def sync_gradients(self) -> float:
        """Sum the flat gradient over the data-parallel ranks; returns the scale Adam must apply."""
        world = self._world()
        if world == 1:
            return 1.0
        if self._overlap is not None and self._overlap["fired"]:
            ov = self._overlap
            if ov["fired"] != len(ov["groups"]) + 1:
                raise MriB200Error("overlapped all-reduce: exactly one backward per optimiser step is required "
                                   "(use overlap=False with gradient accumulation)")
            for work in ov["pending"]:
                work.wait()  # current stream waits for the bucket's all-reduce
            ov["pending"], ov["fired"] = [], 0
            self.allreduce_count += 1
            return (1.0 / world) if self.grad_average else 1.0
        inv_world = allreduce_sum_(self.arena.grad, self.process_group)
        self.allreduce_count += 1
        return inv_world if self.grad_average else 1.0

    # -- overlap of the gradient all-reduce with the hash-grid backward (option 2)
    def enable_overlap(self, encoder, n_groups: int = 4) -> bool:
        """Bucket the table gradients per level group: group g's all-reduce runs on a side stream while the
        scatter kernel of group g+1 executes; the (small) non-table gradients go first.  Requires exactly one
        backward per step.  Returns False (and stays on the single all-reduce) when not applicable."""
        if self._world() == 1 or self.sharded:
            return False
        tables = encoder.tables()
        index = {id(p): i for i, p in enumerate(self.arena.params)}
        if any(id(t) not in index for t in tables):
            return False
        pos = [index[id(t)] for t in tables]
        if pos != list(range(pos[0], pos[0] + len(pos))):
            return False  # tables are not consecutive in the arena
        n_levels = len(tables)
        n_groups = max(1, min(n_groups, n_levels))
        bounds = [round(g * n_levels / n_groups) for g in range(n_groups + 1)]
        groups = [(bounds[g], bounds[g + 1]) for g in range(n_groups) if bounds[g + 1] > bounds[g]]
        offs = self.arena.offsets + [self.arena.numel]
        slices = [(offs[pos[lo]], offs[pos[hi - 1] + 1]) for lo, hi in groups]
        t_begin, t_end = offs[pos[0]], offs[pos[-1] + 1]
        rest = [(0, t_begin), (t_end, self.arena.numel)]
        self._overlap = {"groups": groups, "slices": slices, "rest": [r for r in rest if r[1] > r[0]], "pending": [],
                         "fired": 0, "stream": torch.cuda.Stream(self.arena.grad.device)}
        object.__setattr__(encoder, "_grad_groups", groups)
        object.__setattr__(encoder, "_grad_group_hook", self._on_group_ready)
        return True

    def disable_overlap(self, encoder) -> None:
        if self._overlap is not None:
            for work in self._overlap["pending"]:
                work.wait()
        self._overlap = None
        object.__setattr__(encoder, "_grad_groups", None)
        object.__setattr__(encoder, "_grad_group_hook", None)

    def _on_group_ready(self, gi: int) -> None:
        ov = self._overlap
        if gi == -1 and ov["fired"] != 0:
            raise MriB200Error("overlapped all-reduce: a second backward ran before optimizer.step()")
        spans = ov["rest"] if gi == -1 else [ov["slices"][gi]]
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(ov["stream"]):
            ov["stream"].wait_event(ready)
            for a, b in spans:
                ov["pending"].append(dist.all_reduce(self.arena.grad[a:b], op=dist.ReduceOp.SUM, group=self.process_group,
                                                     async_op=True))
        ov["fired"] += 1

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if not self.arena.intact():
            self.arena.reattach_grads()
            if not self.arena.intact():
                raise MriB200Error("FusedAdam: parameters no longer alias the flat arena (was the model moved "
                                   "after configure_optimizers()?)")
        g = self.param_groups[0]
        if self.sharded and self._world() > 1:
            world = self._world()
            self.step_count += 1
            remote_clear = bool(self._grad_mc) and _MULTICAST_CLEAR
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if self.time_exchange else None
            if ev:
                ev[0].record()
            fused_sync = self._peer_flags is not None
            if not fused_sync:
                self.arena.grad_hdl.barrier(channel=0)  # every rank's backward has finished writing its gradient arena
            if ev:
                ev[1].record()
            args = (world, dist.get_rank(self.process_group),
                    self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.shard_begin, self.shard_len, self.step_count,
                    float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]),
                    (1.0 / world) if self.grad_average else 1.0, 1 if remote_clear else 0, _lib.stream())
            if fused_sync:
                _lib.call("mri_adam_step_sharded_sync", self._peer_grads, self._peer_params, self._grad_mc, self._param_mc,
                          self._peer_flags, *args)
            else:
                _lib.call("mri_adam_step_sharded", self._peer_grads, self._peer_params, self._grad_mc, self._param_mc, *args)
            if ev:
                ev[2].record()
            # new parameters landed everywhere and every slice of my gradient arena has been read by its owner
            if not fused_sync:
                self.arena.data_hdl.barrier(channel=1)
            if ev:
                ev[3].record()
            if not remote_clear:
                # clearing remotely doubles the NVLink stores (P2P pointers: measured 1.265 vs 1.221 ms/step at W=2;
                # multicast: every rank would receive a second arena's worth of zeros), so each rank clears its own arena
                self.arena.grad.zero_()
            if ev:
                ev[4].record()
                ev[4].synchronize()
                self.exchange_times.append(tuple(ev[i].elapsed_time(ev[i + 1]) for i in range(4)))
            self.allreduce_count += 1
            self._grads_clean = True
            return loss
        if self.sharded:
            raise MriB200Error("FusedAdam: a sharded optimiser cannot step with data_parallel switched off")
        scale = self.sync_gradients()
        self.step_count += 1
        if self._step_dev is not None:
            # CUDA-graph friendly: the counter is advanced on the device by the launch itself (replayable)
            _lib.call("mri_adam_step_captured", self.arena.data.data_ptr(), self.arena.grad.data_ptr(),
                      self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.arena.numel, self._step_dev.data_ptr(),
                      self._hyper_dev.data_ptr(), float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
                      float(g["weight_decay"]), float(scale), 1 if self.fuse_zero_grad else 0, _lib.stream(), kernels=2)
        else:
            _lib.call("mri_adam_step", self.arena.data.data_ptr(), self.arena.grad.data_ptr(), self.exp_avg.data_ptr(),
                      self.exp_avg_sq.data_ptr(), self.arena.numel, self.step_count, float(g["lr"]), float(g["betas"][0]),
                      float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]), float(scale),
                      1 if self.fuse_zero_grad else 0, _lib.stream())
        self._grads_clean = bool(self.fuse_zero_grad)
        return loss

    # -- CUDA-graph support (single GPU): see graph.GraphedTrainStep
    def use_device_step(self) -> None:
        """Keep the step counter in device memory from now on, so that step() can be captured in a CUDA graph and
        replayed (host-side bias corrections would be frozen into the graph).  lr/betas/eps are captured by value."""
        if self._world() > 1:
            raise MriB200Error("FusedAdam: CUDA-graph capture of the optimiser step is single-GPU only")
        if self._step_dev is None:
            dev = self.arena.data.device
            self._step_dev = torch.full((1,), self.step_count, dtype=torch.int64, device=dev)
            self._hyper_dev = torch.zeros(2, dtype=torch.float32, device=dev)

    def note_replayed_step(self) -> None:
        """A captured step() was replayed: keep the host-side mirror of the step counter in sync."""
        self.step_count += 1
        self._grads_clean = bool(self.fuse_zero_grad)

    def snapshot(self):
        """Clones of everything a step mutates (parameters, gradients, moments, counter)."""
        return (self.arena.data.clone(), self.arena.grad.clone(), self.exp_avg.clone(), self.exp_avg_sq.clone(),
                self.step_count, self._grads_clean)

    def restore(self, snap) -> None:
        data, grad, m, v, count, clean = snap
        self.arena.data.copy_(data)
        self.arena.grad.copy_(grad)
        self.exp_avg.copy_(m)
        self.exp_avg_sq.copy_(v)
        self.step_count, self._grads_clean = count, clean
        if self._step_dev is not None:
            self._step_dev.fill_(count)

    def zero_grad(self, set_to_none: bool = False):
        """Gradients stay views of the arena.  zero_grad() is free while the arena is known to be all zeros (right after a
        fused step, which clears it, or after another zero_grad()) and NO backward has accumulated into it since -
        e.g. ``step(); backward(); zero_grad()`` does clear the new gradients."""
        intact = self.arena.intact()
        if not self._grads_clean or not intact:
            self.arena.grad.zero_()
        if not intact:
            for p in self.arena.params:  # stale private gradients are dropped, not copied back into the cleared arena
                p.grad = None
            self.arena.reattach_grads()
        self._grads_clean = True

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq,
                "param_groups": [{k: v for k, v in self.param_groups[0].items() if k != "params"}]}

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        if self._step_dev is not None:
            self._step_dev.fill_(self.step_count)
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        for k, v in sd["param_groups"][0].items():
            self.param_groups[0][k] = v
