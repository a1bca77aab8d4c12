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
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

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
        self._staged_for = None
        self._staging_host: Optional[dict] = None
        self._ext_grid: Optional[Tuple[torch.Tensor, torch.Tensor]] = None

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
        if getattr(self.lm, "mask_ratio", 0) != 0:
            # the masked-auto-encoder blocks are drawn on the host for every model call
            # (lightning.mask_tensor): a replayed graph would freeze the first draw
            raise ValueError("mask_ratio != 0 draws a new host-side mask per call: use step(), not capture()/replay()")
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

    # ------------------------------------------------------------- input pipeline
    def configure_external_forcings(self, lat: torch.Tensor, lon: torch.Tensor) -> None:
        """Declares that the LAST five forcing channels are the external forcings of reference
        datasets/base.py:503-515 ([cos_hour, sin_hour, cos_doy, sin_doy, toa_radiation], appended after the
        NWP forcing parameters): `pack_host_batch` then ships their 28-byte-per-slot scalars instead of the
        fields and `load_inputs` rebuilds them on the device.  lat / lon: CUDA tensors (lat, lon), degrees."""
        self._ext_grid = (lat.to(torch.float32).contiguous(), lon.to(torch.float32).contiguous())

    def _ship_dtype(self) -> torch.dtype:
        precision = getattr(getattr(self.lm, "trainer", None), "precision", "32-true")
        return torch.bfloat16 if precision == "bf16-true" else torch.float32

    def pack_host_batch(self, batch: ItemBatch, dates: Optional[Sequence] = None,
                        timedeltas: Optional[Sequence] = None) -> HostBatch:
        """The collate stage of the input pipeline: a host ItemBatch (fp32, every channel, reference
        base.py:173-195) -> pinned tensors in the form that crosses PCIe (see HostBatch).  With
        `configure_external_forcings` and `dates` / `timedeltas` (reference date of every sample, lead times of
        the T steps) the five external channels are dropped and replaced by their scalars."""
        from .forcingutils import external_forcing_scalars

        f = batch.forcing.tensor
        ext = None
        if self._ext_grid is not None and dates is not None:
            ext = external_forcing_scalars(dates, timedeltas).pin_memory()
            f = f[..., : f.shape[-1] - 5]
        pin = lambda t, dt: t.detach().to("cpu", dt).contiguous().pin_memory()  # noqa: E731
        return HostBatch(inputs=pin(batch.inputs.tensor, torch.float32), outputs=pin(batch.outputs.tensor, torch.float32),
                         forcing=pin(f, self._ship_dtype()), ext_scalars=ext)

    def _load_host_batch(self, hb: HostBatch, staged: Optional[dict] = None) -> None:
        """HostBatch (or its staged device copy) -> the captured fp32 input buffers."""
        src = staged if staged is not None else {"inputs": hb.inputs, "outputs": hb.outputs, "forcing": hb.forcing,
                                                 "ext": hb.ext_scalars}
        sb = self._static_batch
        sb.inputs.tensor.copy_(src["inputs"].reshape(sb.inputs.tensor.shape), non_blocking=True)
        sb.outputs.tensor.copy_(src["outputs"].reshape(sb.outputs.tensor.shape), non_blocking=True)
        dst = sb.forcing.tensor
        fs = src["forcing"]
        n_ship = fs.shape[-1]
        if src["ext"] is not None and fs.is_cuda and n_ship + 5 == dst.shape[-1]:
            # staged batch: cast + channel placement + the five generated channels in one pass over whole rows
            from .forcingutils import assemble_forcings

            assemble_forcings(dst, fs.contiguous(), self._ext_grid[0], self._ext_grid[1], src["ext"])
            return
        dst[..., :n_ship].copy_(fs.reshape(*dst.shape[:-1], n_ship), non_blocking=True)  # cast + channel slice
        if src["ext"] is not None:
            from .forcingutils import fill_external_forcings_from_scalars

            ext = src["ext"].to(dst.device, non_blocking=True)
            fill_external_forcings_from_scalars(dst, n_ship, self._ext_grid[0], self._ext_grid[1], ext)
        elif n_ship != dst.shape[-1]:
            raise ValueError("HostBatch ships fewer forcing channels than the model reads and carries no ext_scalars")

    def load_inputs(self, batch) -> None:
        """Copies a batch into the captured input buffers: an ItemBatch (host or device, fp32, every channel)
        or a HostBatch (the pinned, slimmed form of `pack_host_batch`)."""
        if isinstance(batch, HostBatch):
            return self._load_host_batch(batch)
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
            self._staged_ready = torch.cuda.Event()
            self._staging_free = torch.cuda.Event()
            self._staging_free.record(torch.cuda.current_stream())
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._staging_free)  # the previous staged batch was consumed
            if isinstance(batch, HostBatch):
                # same bytes on the device as on the wire: the cast / channel placement / forcing kernel run
                # on the compute stream when the batch is consumed
                dev = self._static_batch.inputs.tensor.device
                if self._staging_host is None or any(
                        self._staging_host[k].shape != t.shape or self._staging_host[k].dtype != t.dtype
                        for k, t in (("inputs", batch.inputs), ("outputs", batch.outputs), ("forcing", batch.forcing))):
                    self._staging_host = {k: torch.empty(t.shape, dtype=t.dtype, device=dev)
                                          for k, t in (("inputs", batch.inputs), ("outputs", batch.outputs),
                                                       ("forcing", batch.forcing))}
                for k, t in (("inputs", batch.inputs), ("outputs", batch.outputs), ("forcing", batch.forcing)):
                    self._staging_host[k].copy_(t, non_blocking=True)
                self._staging_host["ext"] = (None if batch.ext_scalars is None
                                             else batch.ext_scalars.to(dev, non_blocking=True))
            else:
                if self._staging is None:
                    self._staging = _clone_batch(self._static_batch)
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
                if isinstance(batch, HostBatch):
                    self._load_host_batch(batch, staged=self._staging_host)
                else:
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


@dataclass
class HostBatch:
    """What the collate stage hands to `DataParallelTrainer.replay` for one step, in pinned host memory and in
    the form that crosses PCIe (reference: datasets/base.py:173-195 `collate_fn` builds fp32 tensors of every
    channel and Lightning moves them as they are):
      inputs / outputs  fp32 (B, T, lat, lon, F): state and truth stay fp32 (loss and AR state are fp32);
      forcing           the channels that cannot be generated, in the dtype the model input is built in
                        (bf16 under `bf16-true`: the forcing only ever feeds the bf16 `x`, so rounding it on
                        the host is bit-identical to rounding it in the concat kernel);
      ext_scalars       (B * T, 7) fp32 from `forcingutils.external_forcing_scalars`, or None: the five
                        date / solar channels are then rebuilt on the device by p4c_forcing_fill instead of
                        being shipped as per-pixel fields."""

    inputs: torch.Tensor
    outputs: torch.Tensor
    forcing: torch.Tensor
    ext_scalars: Optional[torch.Tensor] = None

    def nbytes(self) -> int:
        n = sum(t.numel() * t.element_size() for t in (self.inputs, self.outputs, self.forcing))
        return n + (self.ext_scalars.numel() * 4 if self.ext_scalars is not None else 0)


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
