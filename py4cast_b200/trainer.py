"""
Batch-sharded data-parallel training step around AutoRegressiveLightning.

Replaces what the reference gets from lightning.Trainer + DDP + torch.optim.AdamW
(config/CLI/trainer.yaml:9,58,62-64; py4cast/lightning.py:442-467, 806-831):
one process per GPU, the graph and the weights replicated, the batch split by rank, ONE
gradient all-reduce per optimiser step.  B200-specific choices:

  * all parameters live in one flat fp32 buffer, all gradients in another: the backward kernels
    accumulate weight gradients straight into it (`param._p4c_grad` sinks, see ops._grad_target),
    the all-reduce is a single NCCL call on that buffer (~5.7 MB for HiLAM: latency-bound over
    NVLink 5 / NVSwitch) and AdamW is a single kernel over it (p4c_adamw_step);
  * the learning-rate schedule (cosine with min-lr and warm-up, the reference's
    transformers.get_cosine_with_min_lr_schedule_with_warmup) advances on the device
    (p4c_adamw_advance), so a whole step can be replayed as a CUDA graph;
  * `capture()` records forward + loss + backward + all-reduce + AdamW into a torch.cuda.CUDAGraph
    (HiLAM issues ~1.5k small launches per 4-step rollout; replay removes the launch latency).
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Callable, List, Optional

import torch
import torch.distributed as dist

from . import _lib, ops
from ._lib import AdamWState, check, ptr, stream
from .datasets import ItemBatch
from .namedtensor import NamedTensor


class FlatParams:
    """Re-homes every parameter of `module` into one flat fp32 buffer (+ a flat grad buffer)."""

    def __init__(self, module: torch.nn.Module):
        params = [p for p in module.parameters() if p.requires_grad]
        if not params:
            raise ValueError("module has no trainable parameter")
        dev = params[0].device
        for p in params:
            if p.dtype != torch.float32 or p.device != dev:
                raise TypeError("FlatParams expects float32 parameters on one device")
        # 4-element (16 B) alignment of every slice
        offs, n = [], 0
        for p in params:
            offs.append(n)
            n += (p.numel() + 3) // 4 * 4
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.params = params
        self.sinks_for_kernels: List[torch.Tensor] = []
        with torch.no_grad():
            for p, o in zip(params, offs):
                view = self.flat[o: o + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view
                gview = self.grad[o: o + p.numel()].view_as(p)
                p.grad = gview
                self.sinks_for_kernels.append(gview)
        self.numel = n

    def enable_kernel_sinks(self, on: bool = True) -> None:
        """Let the backward kernels accumulate into the flat grad buffer directly."""
        for p, g in zip(self.params, self.sinks_for_kernels):
            if on:
                p._p4c_grad = g
            elif hasattr(p, "_p4c_grad"):
                del p._p4c_grad

    def zero_grad(self) -> None:
        self.grad.zero_()


class FusedAdamW:
    """AdamW (torch defaults: eps 1e-8, weight_decay 1e-2) + cosine/min-lr/warm-up schedule."""

    def __init__(self, flat: FlatParams, lr: float, min_lr: float, warmup_steps: int, total_steps: int,
                 betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2):
        self.fp = flat
        self.lr, self.min_lr, self.warmup, self.total = lr, min_lr, warmup_steps, total_steps
        self.betas, self.eps, self.wd = betas, eps, weight_decay
        dev = flat.flat.device
        self.exp_avg = torch.zeros_like(flat.flat)
        self.exp_avg_sq = torch.zeros_like(flat.flat)
        self.state = torch.zeros(C.sizeof(AdamWState), dtype=torch.uint8, device=dev)

    def step(self, grad_scale: float = 1.0) -> None:
        lib = _lib.load()
        check(
            lib.p4c_adamw_advance(ptr(self.state), self.lr, self.min_lr, self.warmup, self.total,
                                  self.betas[0], self.betas[1], stream()),
            "p4c_adamw_advance",
        )
        check(
            lib.p4c_adamw_step(ptr(self.fp.flat), ptr(self.fp.grad), ptr(self.exp_avg), ptr(self.exp_avg_sq),
                               self.fp.numel, ptr(self.state), self.betas[0], self.betas[1], self.eps,
                               self.wd, grad_scale, stream()),
            "p4c_adamw_step",
        )

    def host_state(self) -> dict:
        raw = bytes(self.state.cpu().numpy().tobytes())
        s = AdamWState.from_buffer_copy(raw)
        return {"step": s.step, "lr": s.lr, "bias1": s.bias1, "bias2": s.bias2}


class DataParallelTrainer:
    """One rank of the data-parallel job (world size 1 works without torch.distributed)."""

    def __init__(self, lm, total_steps: int = 1000, use_kernel_sinks: bool = True):
        self.lm = lm
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.flat = FlatParams(lm)
        if self.world > 1:  # identical initial weights on every rank (DDP broadcast at construction)
            dist.broadcast(self.flat.flat, src=0)
        self.flat.enable_kernel_sinks(use_kernel_sinks)
        self.opt = FusedAdamW(
            self.flat, lm.learning_rate, lm.min_learning_rate, lm.num_warmup_steps, total_steps, betas=lm.betas
        )
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._static_batch: Optional[ItemBatch] = None
        self._static_loss: Optional[torch.Tensor] = None
        self._copy_stream: Optional[torch.cuda.Stream] = None
        self._staging: Optional[ItemBatch] = None
        self._staged_for: Optional[ItemBatch] = None

    def sync_gradients(self) -> float:
        """ONE all-reduce (sum) of the flat gradient buffer; returns the scale that turns the sum
        into DDP's mean over ranks (applied inside the AdamW kernel, not as an extra pass)."""
        if self.world > 1:
            dist.all_reduce(self.flat.grad, op=dist.ReduceOp.SUM)
        return 1.0 / self.world

    def _forward_backward(self, batch: ItemBatch) -> torch.Tensor:
        self.flat.zero_grad()
        # stale deferred weight-gradient partials (a backward outside the trainer, or a step that
        # raised before its flush) must not be accumulated into this step
        ops.reset_deferred_wgrads()
        # the reference records feature / dim names on the first batch of every epoch
        # (lightning.py:541-545); `forward()` / phase="inference" need them after training
        batch_idx = 0 if not hasattr(self.lm, "output_feature_names") else 1
        loss = self.lm.training_step(batch, batch_idx)
        self.lm.training_step_losses.clear()
        try:
            loss.backward()
            ops.flush_wgrads()  # batched reduction of the deferred weight-gradient partials
        except BaseException:
            ops.reset_deferred_wgrads()
            raise
        return loss.detach()

    # one optimiser step on `batch` (this rank's shard)
    def _step_impl(self, batch: ItemBatch) -> torch.Tensor:
        loss = self._forward_backward(batch)
        self.opt.step(grad_scale=self.sync_gradients())
        return loss

    def train_step(self, batch: ItemBatch) -> torch.Tensor:
        return self._step_impl(_fresh(batch))

    # ----------------------------------------------------------------- CUDA graph
    def capture(self, example: ItemBatch, warmup: int = 1) -> None:
        """Records zero-grad + forward + loss + backward of one step into a CUDA graph (the
        gradient all-reduce and the 2-kernel AdamW step stay eager, after the replay).
        `example` fixes the shapes; `warmup` full steps are run on it first (lazy set-up work —
        CSR construction, shared-memory opt-in, casts — must not happen under capture)."""
        if warmup < 1:
            raise ValueError("at least one warm-up step is needed before capture")
        self._static_batch = _clone_batch(example)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._step_impl(_fresh(self._static_batch))
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._static_loss = self._forward_backward(_fresh(self._static_batch))

    def load_inputs(self, batch: ItemBatch) -> None:
        """Copies a (host or device) batch into the captured input buffers (H2D when pinned)."""
        for name in ("inputs", "forcing", "outputs"):
            dst = getattr(self._static_batch, name).tensor
            src = getattr(batch, name).tensor
            dst.copy_(src.reshape(dst.shape), non_blocking=True)

    def prefetch(self, batch: ItemBatch) -> None:
        """Starts the host -> device copy of the NEXT step's batch on a dedicated copy stream into
        a staging buffer, so that it overlaps the current step's kernels (what the reference gets
        from DataLoader(pin_memory=True) + non_blocking transfers, lightning.py:572-579).  The
        following `replay(batch)` with the same object moves it into the captured input buffers
        with a device-to-device copy (~0.2 ms for the Titan batch instead of ~10 ms over PCIe)."""
        if self._static_batch is None:
            raise RuntimeError("call capture() first")
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
            self._staging = _clone_batch(self._static_batch)
            self._staged_ready = torch.cuda.Event()
            self._staging_free = torch.cuda.Event()
            self._staging_free.record(torch.cuda.current_stream())
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._staging_free)  # the previous staged batch was consumed
            for name in ("inputs", "forcing", "outputs"):
                dst = getattr(self._staging, name).tensor
                src = getattr(batch, name).tensor
                dst.copy_(src.reshape(dst.shape), non_blocking=True)
            self._staged_ready.record(self._copy_stream)
        self._staged_for = batch

    def replay(self, batch: Optional[ItemBatch] = None, prefetch: Optional[ItemBatch] = None) -> torch.Tensor:
        """One optimiser step: graph replay (fwd+bwd) + all-reduce + AdamW.  `batch` (host or
        device) becomes the step's input — for free when it was announced with `prefetch()`;
        `prefetch` announces the batch of the following step."""
        if self._graph is None:
            raise RuntimeError("call capture() first")
        if batch is not None:
            if self._staged_for is batch:
                cur = torch.cuda.current_stream()
                cur.wait_event(self._staged_ready)
                self.load_inputs(self._staging)
                self._staging_free.record(cur)
                self._staged_for = None
            else:
                self.load_inputs(batch)
        if prefetch is not None:
            self.prefetch(prefetch)
        self._graph.replay()
        self.opt.step(grad_scale=self.sync_gradients())
        return self._static_loss


def shard_indices(num_samples: int, rank: int, world: int) -> list:
    """Sample i goes to rank i mod world (torch DistributedSampler without shuffling/padding)."""
    return list(range(rank, num_samples, world))


def cosine_min_lr_lambda(step: int, warmup_steps: int, total_steps: int, min_lr_rate: float) -> float:
    """Host restatement of the schedule p4c_adamw_advance evaluates on the device
    (transformers.get_cosine_with_min_lr_schedule_with_warmup, num_cycles = 0.5)."""
    import math

    if step < warmup_steps:
        return float(step) / float(max(1, warmup_steps))
    progress = float(step - warmup_steps) / float(max(1, total_steps - warmup_steps))
    factor = 0.5 * (1.0 + math.cos(math.pi * progress))
    return max(0.0, factor * (1 - min_lr_rate) + min_lr_rate)


def _fresh(batch: ItemBatch) -> ItemBatch:
    """New NamedTensor wrappers around the same storage (common_step flattens them in place)."""
    def cp(nt):
        return NamedTensor(nt.tensor, list(nt.names), list(nt.feature_names)) if nt is not None else None

    return ItemBatch(inputs=cp(batch.inputs), forcing=cp(batch.forcing), outputs=cp(batch.outputs))


def _slice_batch(batch: ItemBatch, lo: int, hi: int) -> ItemBatch:
    def cp(nt):
        return NamedTensor(nt.tensor[lo:hi], list(nt.names), list(nt.feature_names)) if nt is not None else None

    return ItemBatch(inputs=cp(batch.inputs), forcing=cp(batch.forcing), outputs=cp(batch.outputs))


def _clone_batch(batch: ItemBatch) -> ItemBatch:
    def cp(nt):
        return NamedTensor(nt.tensor.clone(), list(nt.names), list(nt.feature_names))

    return ItemBatch(inputs=cp(batch.inputs), forcing=cp(batch.forcing), outputs=cp(batch.outputs))


def bind_to_gpu_numa_node(local_rank: int) -> bool:
    """Pins the calling process to the CPUs NVML reports as local to its GPU (same NUMA node / PCIe
    root), so that the pinned host batches allocated afterwards are NUMA-local to the GPU they are
    copied to.  With 8 ranks per box the host -> device copies of the step's inputs (275 MB per sample)
    otherwise cross the socket interconnect (measured at 8 x B200: 22.6 GB/s per GPU unbound).  Best effort: returns False when NVML is unavailable."""
    try:
        import pynvml

        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
        if not uuid.startswith("GPU-"):
            uuid = "GPU-" + uuid
        handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return True
    except Exception:
        return False


def init_distributed(backend: Optional[str] = None) -> tuple:
    """Reads RANK / LOCAL_RANK / WORLD_SIZE (torchrun); returns (rank, local_rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            bind_to_gpu_numa_node(local_rank)
        kw = {"device_id": torch.device("cuda", local_rank)} if backend == "nccl" else {}
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    return rank, local_rank, world
