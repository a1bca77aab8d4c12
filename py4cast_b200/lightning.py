"""
AutoRegressiveLightning: the AR rollout + state update + loss call of the reference, on the
B200 kernels, without the Lightning runtime.

Mirrors reference py4cast/lightning.py:147-306 (construction: feature counting, model registry,
statics / mask buffers, loss preparation), :479-493 (`common_step` autocast wrapper), :495-676
(`_common_step`), :678-709 (`_strategy_params`, `_step_diffs`), :711-767 (`_next_x`), :787-797
(`get_mask_on_nan`), :806-831 (`training_step`) and :442-467 (`configure_optimizers`) — same
method names, arguments, return contract `(pred: NamedTensor, target: NamedTensor)` and the same
in-place flattening of the batch for GRAPH models.

What changed underneath: the per-step `.clone()`, `select`+`cat`,
scaled/diff update, border blend, prev-state roll and the final `torch.stack` are ONE kernel
per AR step (`p4c_ar_step_forward`) writing the new state straight into its slot of the
(B,T,N,F) prediction slab and producing the next model input in the same pass; diff stats are
uploaded once, not every step.

Validation / test (`validation_step`, `test_step`, reference lightning.py:864-917, 995-1063): the
same rollout under `no_grad` on the same kernels, the per-timestep loss, the `ScaledLoss`
MAE / RMSE per-feature metrics the reference computes inside `StateErrorPlot.update`
(plots.py:510-533) and the `reduce_spatial_dim=False` loss map of `SpatialErrorPlot.update`
(plots.py:598-614) — as device tensors kept on the module instead of figures / logger calls.

`mask_on_nan` (:568-569, :600-601, :732-757, :787-797) and `mask_ratio` (:588-589, :769-785) are
served on the same kernels: the NaN masks are taken once per batch, the mask channel rides with
the forcing channels of each model call, the masked-auto-encoder blocks are drawn with the
reference's own `torch.randperm` stream.
`downscaling_only` (experimental upstream; :524, :546-558, :611-621, :759-765): x = cat(statics,
forcings) through the same concat kernel with zero previous states, state = coarse forcing features + y
through the state-update kernel.
Out of scope: figures / loggers / GRIB writing.
"""

from __future__ import annotations

import ctypes as C
from copy import deepcopy
from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from . import _lib, ops
from ._lib import MAX_PREV, ArX, check, dtype_code, ptr, stream
from .base import ModelType, expand_to_batch, features_last_to_second, features_second_to_last
from .datasets import DatasetInfo, ItemBatch, Statics
from .losses import CombinedLoss, ScaledLoss, WeightedLoss  # noqa: F401  (same import surface)
from .models import build_model_from_settings, get_model_kls_and_settings
from .models import registry as model_registry  # noqa: F401
from .namedtensor import NamedTensor

str_to_dtype = {  # reference py4cast/utils.py:104-109
    "bf16-true": torch.bfloat16,
    "16-true": torch.float16,
    "32-true": torch.float32,
    "64-true": torch.float64,
}


class _TrainerStub:
    """Carries `precision` where the reference reads `self.trainer.precision`."""

    def __init__(self, precision: str):
        self.precision = precision
        self.estimated_stepping_batches = 1


def _bstride(t: torch.Tensor) -> int:
    """batch stride (elements) of a (B, N, F) tensor whose (N, F) plane is contiguous"""
    assert t.dim() == 3 and t.stride(2) == 1 and t.stride(1) == t.shape[2], (
        f"expected an (N,F)-contiguous tensor, got shape {tuple(t.shape)} strides {t.stride()}"
    )
    return t.stride(0)


def _fill_arx(a: ArX, prevs: List[Optional[torch.Tensor]], static: torch.Tensor, forcing: torch.Tensor,
              B: int, N: int, F: int) -> None:
    a.batch, a.nodes, a.f_state = B, N, F
    a.f_static, a.f_forcing, a.num_prev = static.shape[-1], forcing.shape[-1], len(prevs)
    for t, p in enumerate(prevs):
        a.prev[t] = ptr(p) if p is not None else None
        a.prev_batch_stride[t] = _bstride(p) if p is not None else 0
    a.stat = ptr(static)
    a.stat_batch_stride = 0 if static.dim() == 2 else _bstride(static)
    a.forcing = ptr(forcing)
    a.forcing_batch_stride = _bstride(forcing)
    a.x_ld = 0


def _new_x(B: int, N: int, f_in: int, x_dtype: torch.dtype, device, a: ArX) -> torch.Tensor:
    """(B,N,f_in) model input whose rows are padded for the tensor-core MLP (bf16: to a multiple of
    8 elements; the AR kernels write zeros into the padding)."""
    x = ops.empty_rows((B, N, f_in), x_dtype, device)
    a.x_ld = x.stride(1)
    return x


def ar_concat(prevs: List[torch.Tensor], static: torch.Tensor, forcing: torch.Tensor,
              x_dtype: torch.dtype) -> torch.Tensor:
    """x = cat(prev[0..T_in-1], static, forcing) along features (reference `_next_x`); `prevs` is
    empty for the downscaling strategy (statics and forcings only)."""
    B, N = forcing.shape[0], forcing.shape[1]
    F = prevs[0].shape[-1] if prevs else 1
    if len(prevs) > MAX_PREV:
        raise ValueError(f"at most {MAX_PREV} input steps are supported")
    f_in = len(prevs) * F + static.shape[-1] + forcing.shape[-1]
    a = ArX()
    _fill_arx(a, prevs, static, forcing, B, N, F)
    x = _new_x(B, N, f_in, x_dtype, forcing.device, a)
    check(_lib.load().p4c_ar_concat(C.byref(a), ptr(x), dtype_code(x_dtype), stream()), "p4c_ar_concat")
    return x


class _ARStepFn(torch.autograd.Function):
    """(y, prev[0..T_in-1]) -> (new_state, x_next)."""

    @staticmethod
    def forward(ctx, cfg: dict, y, *prevs):
        lib = _lib.load()
        ctx.set_materialize_grads(False)
        y = y.contiguous()
        B, N, F = y.shape
        prev_last = prevs[-1]
        truth, border = cfg["truth"], cfg["border"]
        std, mean = cfg["std"], cfg["mean"]
        new_state = cfg["new_state"]  # (B,N,F) view (slab slot or temporary)
        has_next = cfg["forcing_next"] is not None
        x_next, a = None, None
        if has_next:
            static, forcing = cfg["static"], cfg["forcing_next"]
            x_dtype = cfg["x_dtype"]
            f_in = len(prevs) * F + static.shape[-1] + forcing.shape[-1]
            a = ArX()
            _fill_arx(a, list(prevs[1:]) + [None], static, forcing, B, N, F)
            x_next = _new_x(B, N, f_in, x_dtype, y.device, a)
            ctx.f_in, ctx.x_dtype = f_in, x_dtype
        check(
            lib.p4c_ar_step_forward(
                ptr(y), dtype_code(y.dtype), ptr(prev_last), _bstride(prev_last),
                ptr(truth), _bstride(truth) if truth is not None else 0, ptr(border),
                ptr(std), ptr(mean), B, N, F, ptr(new_state), _bstride(new_state),
                C.byref(a) if has_next else None, ptr(x_next),
                dtype_code(x_next.dtype) if has_next else 0, stream(),
            ),
            "p4c_ar_step_forward",
        )
        ctx.border, ctx.std = border, std
        # gradient hand-over between consecutive steps of a slab rollout (see backward)
        ctx.holder, ctx.slot, ctx.prev_last_slot = cfg.get("holder"), cfg.get("slot"), cfg.get("prev_last_slot")
        ctx.dims = (B, N, F, len(prevs))
        ctx.y_dtype = y.dtype
        ctx.has_next = has_next
        if has_next:
            return new_state, x_next
        return new_state

    @staticmethod
    def backward(ctx, g_new, g_x=None):
        lib = _lib.load()
        B, N, F, T = ctx.dims
        dev = ctx.border.device if ctx.border is not None else (g_new if g_new is not None else g_x).device
        # The state this step produced feeds the loss (through the slab) AND the next step.  Instead of letting
        # autograd add the two gradients with an element-wise pass (a strided slab view + a dense tensor), the
        # next step's backward parks its input gradient in the rollout's holder and this kernel adds it.
        holder = ctx.holder
        g_new2 = None
        if holder is not None and ctx.slot is not None:
            g_new2 = holder["carry"].pop(ctx.slot, None)
        if g_new is None and g_new2 is not None:
            g_new, g_new2 = g_new2, None
        if g_new is None and g_x is None:
            return (None,) * (2 + T)
        if g_new is not None:
            if g_new.dtype != torch.float32:
                g_new = g_new.float()
            if not (g_new.stride(2) == 1 and g_new.stride(1) == F):
                g_new = g_new.contiguous()
        x_ld = 0
        if g_x is not None:
            # rows may be padded (the gradient of a padded x comes back with the same row stride)
            if not (g_x.stride(2) == 1 and g_x.stride(1) >= ctx.f_in
                    and (B == 1 or g_x.stride(0) == N * g_x.stride(1))):
                g_x = g_x.contiguous()
            x_ld = g_x.stride(1)
        dy = torch.empty(B, N, F, dtype=ctx.y_dtype, device=dev)
        d_prev: List[Optional[torch.Tensor]] = [None] * T
        arr = (C.c_void_p * T)()
        for t in range(T):
            need = ctx.needs_input_grad[2 + t]
            used = (t == T - 1) or (g_x is not None and t >= 1)
            if need and used:
                d_prev[t] = torch.empty(B, N, F, dtype=torch.float32, device=dev)
                arr[t] = ptr(d_prev[t])
        check(
            lib.p4c_ar_step_backward(
                ptr(g_new), _bstride(g_new) if g_new is not None else 0,
                ptr(g_new2), _bstride(g_new2) if g_new2 is not None else 0, ptr(g_x),
                dtype_code(g_x.dtype) if g_x is not None else 0, ptr(ctx.border), ptr(ctx.std),
                B, N, F, T, x_ld, ptr(dy), dtype_code(ctx.y_dtype),
                arr, stream(),
            ),
            "p4c_ar_step_backward",
        )
        if (holder is not None and holder["slab_seen"] and ctx.prev_last_slot is not None
                and d_prev[T - 1] is not None):
            # prev_last is a slab state whose own backward has not run yet (it precedes this step) and will run
            # (the slab gradient reaches it): it picks this tensor up as g_new2
            holder["carry"][ctx.prev_last_slot] = d_prev[T - 1]
            d_prev[T - 1] = None
        return (None, dy, *d_prev)


class _SlabFn(torch.autograd.Function):
    """Ties the T per-step states (views of the slab) to the (B,T,N,F) prediction tensor
    (replaces torch.stack, reference lightning.py:660-662) without copying."""

    @staticmethod
    def forward(ctx, slab, holder, *states):
        ctx.T = len(states)
        ctx.holder = holder
        return slab.view(slab.shape)

    @staticmethod
    def backward(ctx, g):
        if ctx.holder is not None:
            ctx.holder["slab_seen"] = True  # every state receives a gradient: the steps may hand theirs over
        return (None, None, *[g[:, i] for i in range(ctx.T)])


class AutoRegressiveLightning(nn.Module):
    """Auto-regressive module for predicting meteorological fields (reference lightning.py:147)."""

    def __init__(
        self,
        settings_init_args: dict,
        dataset_info: DatasetInfo,
        infer_ds=None,
        dataset_name: str = "dummy",
        dataset_conf: Optional[Dict] = None,
        num_input_steps: int = 1,
        num_pred_steps_train: int = 1,
        num_pred_steps_val_test: int = 1,
        batch_size: int = 2,
        model_name: str = "HalfUNet",
        losses: Optional[List[dict]] = None,
        num_inter_steps: int = 1,
        num_samples_to_plot: int = 1,
        training_strategy: str = "diff_ar",
        channels_last: bool = False,
        io_conf=None,
        mask_ratio: float = 0,
        mask_on_nan: bool = False,
        learning_rate: float = 1e-4,
        min_learning_rate: float = 1e-6,
        num_warmup_steps: int = 0,
        betas: tuple = (0.9, 0.999),
        precision: str = "32-true",
    ):
        super().__init__()
        if losses is None:
            losses = [{"class": "WeightedLoss", "params": {"loss": "MSELoss", "reduction": "none"}}]
        self.infer_ds = infer_ds
        self.settings_init_args = settings_init_args
        self.dataset_name, self.dataset_conf, self.dataset_info = dataset_name, dataset_conf, dataset_info
        self.batch_size, self.model_name = batch_size, model_name
        self.num_input_steps = num_input_steps
        self.num_pred_steps_train, self.num_pred_steps_val_test = num_pred_steps_train, num_pred_steps_val_test
        self.num_inter_steps, self.num_samples_to_plot = num_inter_steps, num_samples_to_plot
        self.training_strategy, self.channels_last = training_strategy, channels_last
        self.io_conf, self.mask_ratio, self.mask_on_nan = io_conf, mask_ratio, mask_on_nan
        self.learning_rate, self.min_learning_rate = learning_rate, min_learning_rate
        self.num_warmup_steps, self.betas = num_warmup_steps, tuple(betas)
        self.trainer = _TrainerStub(precision)
        if precision not in str_to_dtype:
            raise KeyError(precision)

        if self.num_inter_steps > 1 and self.num_input_steps > 1:
            raise AttributeError(
                "It is not possible to have multiple input steps when num_inter_steps > 1."
                f"Get num_input_steps :{self.num_input_steps} and num_inter_steps: {self.num_inter_steps}"
            )
        ALLOWED_STRATEGIES = ("diff_ar", "scaled_ar", "downscaling_only")
        if self.training_strategy not in ALLOWED_STRATEGIES:
            raise AttributeError(
                f"Unknown strategy {self.training_strategy}, allowed strategies are {ALLOWED_STRATEGIES}"
            )
        if self.training_strategy == "downscaling_only":
            print("WARNING : You are using downscaling_only mode: this is experimental.")

        statics: Statics = deepcopy(dataset_info.statics)
        self.diff_stats = dataset_info.diff_stats
        self.stats = dataset_info.stats
        self.grid_shape = statics.grid_shape
        self.training_step_losses: list = []

        num_grid_static_features = statics.grid_statics.dim_size("features")
        num_input_features = (
            num_input_steps * dataset_info.weather_dim * (self.training_strategy != "downscaling_only")
            + num_grid_static_features + dataset_info.forcing_dim
            + self.mask_on_nan  # the "valid pixel" channel of _next_x (reference lightning.py:256-261)
        )
        num_output_features = dataset_info.weather_dim

        model_kls, model_settings = get_model_kls_and_settings(model_name, self.settings_init_args)
        # reference: rank-0-only graph creation without a barrier (lightning.py:269-271); here every
        # rank builds the (deterministic) graph itself and writes it atomically.
        if hasattr(model_kls, "rank_zero_setup"):
            model_kls.rank_zero_setup(model_settings, statics.meshgrid)
        self.model, model_settings = build_model_from_settings(
            model_name, num_input_features, num_output_features, self.settings_init_args, statics.grid_shape
        )
        digest = getattr(self.model, "graph_digest", None)
        if digest is not None:
            from .nlam import _meshgrid_digest

            if digest != _meshgrid_digest(statics.meshgrid):
                raise ValueError(
                    "the cached mesh graph under settings.tmp_dir was built for another domain "
                    "(same grid shape, other coordinates): use a tmp_dir per domain"
                )
        if channels_last:
            self.model = self.model.to(memory_format=torch.channels_last)

        if self.model.model_type == ModelType.GRAPH:
            statics.grid_statics.flatten_("ngrid", 0, 1)
            statics.border_mask = statics.border_mask.flatten(0, 1)
            statics.interior_mask = statics.interior_mask.flatten(0, 1)
        statics.register_buffers(self)  # border_mask, interior_mask
        self.num_spatial_dims = statics.grid_statics.num_spatial_dims
        self.register_buffer(
            "grid_static_features", expand_to_batch(statics.grid_statics.tensor, batch_size), persistent=False
        )
        # flat (N, F_s) / (N,) fp32 copies consumed by the kernels (shared by all samples)
        self.register_buffer(
            "_static_flat", statics.grid_statics.tensor.reshape(-1, num_grid_static_features).contiguous().float(),
            persistent=False,
        )
        self.register_buffer("_border_flat", statics.border_mask.reshape(-1).contiguous().float(), persistent=False)

        self.loss = CombinedLoss(losses)
        self.loss.prepare(self, statics.interior_mask, dataset_info)
        self._diff_cache: dict = {}

    # ------------------------------------------------------------------ helpers
    @property
    def dtype(self):
        return str_to_dtype[self.trainer.precision]

    def configure_optimizers(self):
        """torch AdamW + cosine-with-min-lr schedule (reference lightning.py:442-467).
        py4cast_b200.trainer uses the fused flat AdamW kernel with the same maths instead."""
        from transformers.optimization import get_cosine_with_min_lr_schedule_with_warmup

        optimizer = torch.optim.AdamW(self.parameters(), lr=self.learning_rate, betas=self.betas)
        scheduler = get_cosine_with_min_lr_schedule_with_warmup(
            optimizer, num_warmup_steps=self.num_warmup_steps,
            num_training_steps=self.trainer.estimated_stepping_batches, min_lr=self.min_learning_rate,
        )
        return {"optimizer": optimizer, "lr_scheduler": {"scheduler": scheduler, "interval": "step", "frequency": 1}}

    def forward(self, x: ItemBatch, batch_idx: int) -> NamedTensor:
        return self.common_step(x, batch_idx, phase="inference")[0]

    def common_step(self, batch: ItemBatch, batch_idx: int, phase: str) -> Tuple[NamedTensor, NamedTensor]:
        """Autocast wrapper (reference lightning.py:479-493)."""
        if self.dtype in (torch.bfloat16, torch.float16):
            with torch.amp.autocast("cuda", dtype=self.dtype):
                return self._common_step(batch, batch_idx, phase)
        return self._common_step(batch, batch_idx, phase)

    def _strategy_params(self) -> Tuple[bool, bool, int]:
        force_border = self.training_strategy == "scaled_ar"
        scale_y = self.training_strategy == "scaled_ar"
        if self.training_strategy == "diff_ar" and self.num_inter_steps != 1:
            raise ValueError("Diff AR strategy requires exactly 1 intermediary step.")
        return force_border, scale_y, self.num_inter_steps

    def _step_diffs(self, feature_names: List[str], device: torch.device) -> Tuple[torch.Tensor, torch.Tensor]:
        """(diff_std, diff_mean) on `device`, by feature-name order; uploaded ONCE and cached
        (the reference re-stacks and re-uploads them every AR step, lightning.py:696-709)."""
        key = (tuple(feature_names), str(device))
        if key not in self._diff_cache:
            std = self.diff_stats.to_list("std", feature_names).to(device).contiguous()
            mean = self.diff_stats.to_list("mean", feature_names).to(device).contiguous()
            self._diff_cache[key] = (std, mean)
        return self._diff_cache[key]

    def _x_dtype(self) -> torch.dtype:
        if self.model.model_type == ModelType.GRAPH and self.dtype == torch.bfloat16:
            return torch.bfloat16
        return torch.float32

    def _next_x(self, batch: ItemBatch, prev_states: NamedTensor, step_idx: int) -> torch.Tensor:
        """First-step model input: cat(prev states, statics, forcing) (reference lightning.py:711-767)."""
        T_in = batch.num_input_steps
        t = prev_states.tensor
        flat = t.reshape(t.shape[0], t.shape[1], -1, t.shape[-1]).float()
        f = batch.forcing.select_tensor_dim("timestep", step_idx)
        forcing = f.reshape(f.shape[0], -1, f.shape[-1]).float()
        if self.mask_on_nan:
            self._require_grid_layout(t.dim() - 1)
            valid = ~(torch.isnan(flat[:, :T_in]).any(-1).any(1) | torch.isnan(forcing).any(-1))
            flat = torch.nan_to_num(flat, nan=0)
            forcing = torch.cat([torch.nan_to_num(forcing, nan=0), valid.unsqueeze(-1).float()], dim=-1)
        prevs = [flat[:, i].contiguous() for i in range(T_in)]
        return ar_concat(prevs, self._static_flat, forcing.contiguous(), self._x_dtype())

    @staticmethod
    def match_common_features(output_names: List[str], forcing_names: List[str]) -> List[int]:
        """Indices, in output-feature order, of the forcing features that are the coarse version of an output
        feature: same name after the first "_"-separated token ("aro_t2m_2m" ~ "arp_t2m_2m"); every match counts,
        as in the reference's double loop (lightning.py:546-558)."""
        return [i for out_name in output_names for i, f_name in enumerate(forcing_names)
                if out_name.split("_")[1:] == f_name.split("_")[1:]]

    @staticmethod
    def _require_grid_layout(ndim: int) -> None:
        """The reference's NaN-mask code indexes (batch, lat, lon, features) tensors
        (lightning.py:733-748): on flattened graph inputs it raises IndexError; so does this."""
        if ndim != 4:
            raise IndexError("too many indices for tensor of dimension 3 (mask_on_nan needs gridded inputs)")

    @staticmethod
    def block_mask(height: int, width: int, mask_ratio: float) -> torch.Tensor:
        """(H, W) bool mask of the masked-auto-encoder strategy, False inside the dropped blocks:
        the blocks `mask_tensor` of the reference clears one slice assignment at a time
        (lightning.py:769-785), drawn from the same `torch.randperm` stream (CPU default
        generator), painted here with two outer products instead of a Python loop."""
        num_blocks = int((1 - mask_ratio) * height * width)
        bh, bw = height // int(height**0.5), width // int(width**0.5)
        idx = torch.randperm(height * width)[:num_blocks]
        rows, cols = torch.div(idx, width, rounding_mode="floor"), idx % width
        # block (row, col) covers [row*bh, (row+1)*bh) x [col*bw, (col+1)*bw), clipped by slicing
        hit = torch.zeros(height // bh + 1, width // bw + 1, dtype=torch.bool)
        inside = (rows * bh < height) & (cols * bw < width)
        hit[rows[inside], cols[inside]] = True
        full = hit.repeat_interleave(bh, 0)[:height].repeat_interleave(bw, 1)[:, :width]
        return ~full

    def mask_tensor(self, x: torch.Tensor) -> torch.Tensor:
        """x * mask (reference lightning.py:769-785); x is (B, H, W, C)."""
        _, height, width, _ = x.shape
        mask = self.block_mask(height, width, self.mask_ratio).to(x.device)
        return x * mask.view(1, height, width, 1)

    def _run_model(self, x_flat: torch.Tensor, spatial: Tuple[int, ...]) -> torch.Tensor:
        """x_flat (B,N,F_in) -> y (B,N,F), through the layout the model type expects."""
        if self.model.model_type == ModelType.GRAPH:
            if self.mask_ratio != 0:  # the reference unpacks four dimensions (lightning.py:770)
                raise ValueError("not enough values to unpack (expected 4, got 3): mask_ratio needs gridded inputs")
            return self.model(x_flat)
        x = x_flat.view(x_flat.shape[0], *spatial, x_flat.shape[-1])
        if self.channels_last:
            x = x.to(memory_format=torch.channels_last)
        if self.mask_ratio != 0:  # masked-auto-encoder strategy (lightning.py:588-589)
            x = self.mask_tensor(x)
        if self.model.features_second:
            y = features_second_to_last(self.model(features_last_to_second(x)))
        else:
            y = self.model(x)
        return y.reshape(y.shape[0], -1, y.shape[-1])

    # ------------------------------------------------------------- the AR rollout
    def _common_step(self, batch: ItemBatch, batch_idx: int, phase: str) -> Tuple[NamedTensor, NamedTensor]:
        force_border, scale_y, num_inter_steps = self._strategy_params()
        self.original_shape = None
        inference = phase == "inference"

        if self.model.model_type == ModelType.GRAPH:
            self.original_shape = batch.inputs.tensor.shape
            batch.inputs.flatten_("ngrid", *batch.inputs.spatial_dim_idx)
            if not inference:
                batch.outputs.flatten_("ngrid", *batch.outputs.spatial_dim_idx)
            batch.forcing.flatten_("ngrid", *batch.forcing.spatial_dim_idx)

        if batch_idx == 0 and batch.outputs is not None:
            self.input_feature_names = batch.inputs.feature_names
            self.output_feature_names = batch.outputs.feature_names
            self.output_dim_names = batch.outputs.names
            self.output_dtype = batch.outputs.tensor.dtype
            if self.training_strategy == "downscaling_only":
                # forcing features that are the coarse version of an output feature: same name after
                # the first "_"-separated token (reference lightning.py:546-558)
                self.common_features_idx = self.match_common_features(
                    self.output_feature_names, batch.forcing.feature_names)

        inp = batch.inputs.tensor
        if not inp.is_cuda:
            raise _lib.P4CError("py4cast_b200 runs on CUDA only (no CPU fallback): move the batch to a cuda device")
        B, T_in, F = inp.shape[0], inp.shape[1], inp.shape[-1]
        spatial = tuple(inp.shape[2:-1])
        N = 1
        for s in spatial:
            N *= s
        T = batch.num_pred_steps
        dev = inp.device
        inputs = inp.reshape(B, T_in, N, F).float().contiguous()
        forcing = batch.forcing.tensor.reshape(B, T, N, -1).float().contiguous()
        outputs = None if inference else batch.outputs.tensor.reshape(B, T, N, F).float().contiguous()
        names = self.output_feature_names if inference else batch.outputs.feature_names
        std, mean = self._step_diffs(names, dev) if scale_y else (None, None)
        border = self._border_flat if (force_border and not inference) else None
        x_dtype = self._x_dtype()
        call_forcing = None
        if self.mask_on_nan:
            # NaN handling of the reference (:568-569 border state, :600-601 last state, :732-757 x):
            # every NaN of inputs / forcing / truth reads as 0 and x gains a "valid pixel" channel =
            # no NaN in any state of the input window nor in the step's forcing.  States produced by
            # the rollout are NaN-free, so only the raw input states still in the window count.
            self._require_grid_layout(len(spatial) + 2)
            in_nan = torch.isnan(inputs).any(-1)  # (B, T_in, N)
            f_nan = torch.isnan(forcing).any(-1)  # (B, T, N)
            inputs, forcing = torch.nan_to_num(inputs, nan=0), torch.nan_to_num(forcing, nan=0)
            if outputs is not None:
                outputs = torch.nan_to_num(outputs, nan=0)
            call_forcing = torch.empty(B, T * num_inter_steps, N, forcing.shape[-1] + 1, dtype=torch.float32, device=dev)
            for c in range(T * num_inter_steps):
                invalid = f_nan[:, c // num_inter_steps]
                if c < T_in:
                    invalid = invalid | in_nan[:, c:].any(1)
                call_forcing[:, c, :, :-1] = forcing[:, c // num_inter_steps]
                call_forcing[:, c, :, -1] = (~invalid).float()

        def forcing_of_call(c: int) -> torch.Tensor:
            """forcing channels (+ mask channel) of model call c = i * num_inter_steps + k"""
            return call_forcing[:, c] if call_forcing is not None else forcing[:, c // num_inter_steps]

        slab = torch.empty(B, T, N, F, dtype=torch.float32, device=dev)
        if hasattr(self.model, "begin_rollout"):
            # batch-independent embeddings and weight images: once per rollout, not per step
            self.model.begin_rollout(x_dtype)
        ds = self.training_strategy == "downscaling_only"
        if ds:
            # residual on the coarse forcing features, no recurrence (reference lightning.py:611-621):
            # the state-update kernel runs with "previous state" = the selected forcing channels
            self._require_grid_layout(len(spatial) + 2)
            idx = torch.as_tensor(self.common_features_idx, dtype=torch.long, device=dev)
            if idx.numel() != F:
                raise RuntimeError(
                    f"The size of tensor a ({idx.numel()}) must match the size of tensor b ({F}) at non-singleton "
                    "dimension 3 (downscaling_only: forcing features matching the output features by name)"
                )
            coarse = forcing.index_select(-1, idx)  # (B, T, N, F)
        try:
            states = []
            holder = {"slab_seen": False, "carry": {}}
            prev_last_slot = None  # slab slot of prevs[-1] when it is a state of this rollout
            prevs: List[torch.Tensor] = [inputs[:, t] for t in range(T_in)]
            if not ds:
                x = ar_concat(prevs, self._static_flat, forcing_of_call(0), x_dtype)
            for i in range(T if ds else 0):
                for k in range(num_inter_steps):
                    x = ar_concat([], self._static_flat, forcing_of_call(i * num_inter_steps + k), x_dtype)
                    y = self._run_model(x, spatial)
                    last_k = k == num_inter_steps - 1
                    cfg = {
                        "truth": None, "border": None, "std": None, "mean": None,
                        "new_state": slab[:, i] if last_k else torch.empty(B, N, F, dtype=torch.float32, device=dev),
                        "static": self._static_flat, "forcing_next": None, "x_dtype": x_dtype,
                        "holder": holder, "slot": i if last_k else None, "prev_last_slot": None,
                    }
                    new_state = _ARStepFn.apply(cfg, y, coarse[:, i])
                states.append(new_state)
            for i in range(0 if ds else T):
                for k in range(num_inter_steps):
                    y = self._run_model(x, spatial)
                    last_k = k == num_inter_steps - 1
                    more = i < T - 1 or not last_k
                    next_forcing = forcing_of_call(i * num_inter_steps + k + 1) if more else None
                    cfg = {
                        "truth": outputs[:, i] if border is not None else None,
                        "border": border, "std": std, "mean": mean,
                        "new_state": slab[:, i] if last_k else torch.empty(B, N, F, dtype=torch.float32, device=dev),
                        "static": self._static_flat, "forcing_next": next_forcing, "x_dtype": x_dtype,
                        "holder": holder, "slot": i if last_k else None, "prev_last_slot": prev_last_slot,
                    }
                    res = _ARStepFn.apply(cfg, y, *prevs)
                    if more:
                        new_state, x = res
                        prevs = prevs[1:] + [new_state]
                    else:
                        new_state = res
                    prev_last_slot = i if last_k else None
                states.append(new_state)
        finally:
            # always paired: an exception inside the rollout must not leave stale embeddings /
            # weight images behind for the next one
            if hasattr(self.model, "end_rollout"):
                self.model.end_rollout()
        prediction = _SlabFn.apply(slab, holder, *states).view(B, T, *spatial, F)

        if inference:
            pred_out = NamedTensor(prediction.type(self.output_dtype), self.output_dim_names, self.output_feature_names)
        else:
            pred_out = NamedTensor.new_like(prediction.type_as(batch.outputs.tensor), batch.outputs)
        return pred_out, batch.outputs

    def get_mask_on_nan(self, target: NamedTensor):
        """(mask, target with NaNs as zeros) (reference lightning.py:787-797).  Without `mask_on_nan`
        `None` stands for the all-ones mask the reference allocates, so the loss kernel never reads it."""
        if self.mask_on_nan:
            mask = ~torch.isnan(target.tensor)
            target_masked = target.clone()
            target_masked.tensor = torch.nan_to_num(target_masked.tensor, nan=0)
            return mask, target_masked
        return None, target

    def training_step(self, batch: ItemBatch, batch_idx: int) -> torch.Tensor:
        """Train on a single batch (reference lightning.py:806-831, plotters excluded)."""
        prediction, target = self.common_step(batch, batch_idx, phase="train")
        mask, target_masked = self.get_mask_on_nan(target)
        batch_loss = torch.mean(self.loss(prediction, target_masked, mask=mask))
        self.training_step_losses.append(batch_loss.detach())
        return batch_loss

    # ------------------------------------------------------- validation / test
    def _state_metrics(self, which: Tuple[Tuple[str, str], ...]) -> Dict[str, ScaledLoss]:
        """`ScaledLoss` metrics by alias, prepared once (reference on_validation_start :864-869
        builds {"mae": L1}, on_test_start :995-1001 builds {"mae": L1, "rmse": MSE})."""
        cache = self.__dict__.setdefault("_metric_cache", {})
        out = {}
        for torch_loss, alias in which:
            if alias not in cache:
                m = ScaledLoss(torch_loss, reduction="none")
                m.prepare(self, self.interior_mask, self.dataset_info)
                cache[alias] = m
            out[alias] = cache[alias]
        return out

    def _eval_step(self, batch: ItemBatch, batch_idx: int, metrics: Dict[str, ScaledLoss],
                   spatial_map: bool) -> Dict[str, torch.Tensor]:
        with torch.no_grad():
            prediction, target = self.common_step(batch, batch_idx, phase="val_test")
            mask, target_masked = self.get_mask_on_nan(target)
            time_step_loss = torch.mean(self.loss(prediction, target_masked, mask), dim=0)
            out = {"time_step_loss": time_step_loss, "mean_loss": torch.mean(time_step_loss)}
            for alias, metric in metrics.items():
                out[alias] = metric(prediction, target_masked, mask)  # (B, T, F)
            if spatial_map:
                spatial_loss = self.loss(prediction, target_masked, mask, reduce_spatial_dim=False)
                if prediction.num_spatial_dims == 1:  # "b t (x y) -> b t x y" (plots.py:608-611)
                    spatial_loss = spatial_loss.reshape(*spatial_loss.shape[:2], *self.grid_shape)
                out["spatial_loss"] = spatial_loss
        out["prediction"], out["target"] = prediction, target_masked
        return out

    def validation_step(self, batch: ItemBatch, batch_idx: int) -> Dict[str, torch.Tensor]:
        """Runs validation on a single batch (reference lightning.py:887-917): rollout without
        autograd, loss per timestep, MAE per feature.  Returns the tensors the reference hands
        to its loggers / plotters; `val_mean_loss` and `validation_step_losses` are kept as there."""
        out = self._eval_step(batch, batch_idx, self._state_metrics((("L1Loss", "mae"),)), spatial_map=False)
        self.__dict__.setdefault("validation_step_losses", []).append(out["mean_loss"])
        self.val_mean_loss = out["mean_loss"]
        return out

    def test_step(self, batch: ItemBatch, batch_idx: int) -> Dict[str, torch.Tensor]:
        """Runs test on a single batch (reference lightning.py:1017-1042 + the plotters'
        device-side work: MAE / RMSE per feature and the spatial loss map)."""
        metrics = self._state_metrics((("L1Loss", "mae"), ("MSELoss", "rmse")))
        return self._eval_step(batch, batch_idx, metrics, spatial_map=True)
