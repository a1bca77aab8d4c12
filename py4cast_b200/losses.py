"""
Losses on the B200 kernels — same class names, constructor arguments, `prepare` / `forward`
interface and return shapes as reference py4cast/losses.py:18-307 (Py4CastLoss, WeightedLoss,
ScaledLoss, CombinedLoss), so the YAML form
    losses: [{class: WeightedLoss, weight: 1.0, params: {loss: MSELoss, reduction: "none"}}]
(config/CLI/model/hilam.yaml:3-8) resolves unchanged.

Differences, on purpose:
  * forward + backward are one reduction kernel + one elementwise kernel
    (p4c_loss_forward / p4c_loss_backward) instead of ~8 full-tensor ATen passes;
  * `mask=None` means "all ones" (reference lightning.py:797 allocates torch.ones_like(target));
    a mask tensor is honoured exactly like the reference (p*m, t*m, union-mask denominator);
  * 4-D graph tensors (B,T,N,F) work: the reference hard-codes the feature dim as `4` in
    `torch.any(mask, dim=(0, 1, 4))` (losses.py:156,197) and raises IndexError on them;
  * PerceptualLossPy4Cast is out of scope (needs mfai's pretrained CNN).
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from functools import lru_cache
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import check, ptr, stream
from .namedtensor import NamedTensor

_KINDS = {"MSELoss": 0, "L1Loss": 1}
FORM_WEIGHTED, FORM_SCALED, FORM_MAP = 0, 1, 2


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, mask, interior, w, num_interior: float, kind: int, form: int):
        lib = _lib.load()
        if pred.dtype != torch.float32 or target.dtype != torch.float32:
            raise TypeError("loss kernels compute in fp32: prediction and target must be float32")
        pred, target = pred.contiguous(), target.contiguous()
        F = pred.shape[-1]
        BT = pred.shape[0] * pred.shape[1]
        N = pred.numel() // (BT * F)
        dev = pred.device
        if mask is not None:
            mask = mask.to(torch.float32).expand_as(pred).contiguous()
        denom = torch.empty(1, dtype=torch.float32, device=dev)
        if form != FORM_MAP:
            scratch = torch.empty(1, dtype=torch.int32, device=dev) if mask is not None else None
            check(
                lib.p4c_loss_denominator(ptr(mask), BT, N, F, float(num_interior), ptr(denom),
                                         ptr(scratch), stream()),
                "p4c_loss_denominator",
            )
        if form == FORM_WEIGHTED:
            out = torch.empty(pred.shape[:2], dtype=torch.float32, device=dev)
        elif form == FORM_SCALED:
            out = torch.empty(*pred.shape[:2], F, dtype=torch.float32, device=dev)
        else:
            out = torch.empty(pred.shape[:-1], dtype=torch.float32, device=dev)
        nfl = lib.p4c_loss_partial_floats(BT, N, F, form)
        partials = torch.empty(max(int(nfl), 1), dtype=torch.float32, device=dev)
        check(
            lib.p4c_loss_forward(ptr(pred), ptr(target), ptr(mask), ptr(interior), ptr(w), ptr(denom),
                                 BT, N, F, kind, form, ptr(partials), ptr(out), stream()),
            "p4c_loss_forward",
        )
        ctx.save_for_backward(pred, target, mask, interior, w, denom, out)
        ctx.dims = (BT, N, F, kind, form)
        return out

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load()
        pred, target, mask, interior, w, denom, out = ctx.saved_tensors
        BT, N, F, kind, form = ctx.dims
        gout = gout.to(torch.float32).contiguous()
        dpred = torch.empty_like(pred)
        dtarget = torch.empty_like(target) if ctx.needs_input_grad[1] else None
        check(
            lib.p4c_loss_backward(ptr(pred), ptr(target), ptr(mask), ptr(interior), ptr(w), ptr(denom),
                                  ptr(gout), ptr(out), BT, N, F, kind, form, ptr(dpred), ptr(dtarget),
                                  stream()),
            "p4c_loss_backward",
        )
        return dpred, dtarget, None, None, None, None, None, None


class Py4CastLoss(ABC):
    """Base class: `loss` is the name of a torch.nn loss class (reference losses.py:25-31)."""

    def __init__(self, loss: str, *args, **kwargs) -> None:
        if not hasattr(torch.nn, loss):
            raise NameError(f"Loss: {loss} is not defined")
        if loss not in _KINDS:
            raise NotImplementedError(
                f"Loss: {loss} has no B200 kernel; supported: {sorted(_KINDS)} (reduction='none')"
            )
        if kwargs.get("reduction", "none") != "none":
            raise ValueError("py4cast losses do their own reduction: pass reduction='none'")
        self.loss_name = loss
        self.kind = _KINDS[loss]
        # kept for API compatibility (`self.loss.__class__ == MSELoss` checks in user code)
        self.loss = getattr(torch.nn, loss)(*args, **kwargs)

    @abstractmethod
    def prepare(self, lm: torch.nn.Module, interior_mask: torch.Tensor, dataset_info) -> None: ...

    @abstractmethod
    def forward(self, prediction: NamedTensor, target: NamedTensor, mask: Optional[torch.Tensor]) -> torch.Tensor: ...

    def register_loss_state_buffers(
        self, lm: torch.nn.Module, interior_mask: torch.Tensor, loss_state_weight: dict,
        squeeze_mask: bool = False,
    ) -> None:
        """Registers the (squeezed) interior mask on the module, keeps the weight dict
        (reference losses.py:52-72)."""
        self.loss_state_weight = loss_state_weight
        attr_name = "interior_mask_s" if squeeze_mask else "interior_mask"
        if not hasattr(lm, attr_name):
            lm.register_buffer(
                attr_name, interior_mask.squeeze(-1) if squeeze_mask else interior_mask, persistent=False
            )
        self.num_interior = torch.sum(interior_mask).item()

    def __call__(self, *args, **kwds):
        return self.forward(*args, **kwds)

    @lru_cache(maxsize=8)
    def weights(self, feature_names: Tuple[str], device: torch.device) -> torch.Tensor:
        return torch.stack(
            [torch.as_tensor(self.loss_state_weight[name], dtype=torch.float32) for name in feature_names]
        ).to(device, non_blocking=True)

    def _interior_flat(self, device) -> torch.Tensor:
        key = str(device)
        cache = self.__dict__.setdefault("_interior_cache", {})
        if key not in cache:
            m = getattr(self.lm, "interior_mask_s", None)
            if m is None:
                m = self.lm.interior_mask.squeeze(-1)
            cache[key] = m.to(device=device, dtype=torch.float32).reshape(-1).contiguous()
        return cache[key]


class WeightedLoss(Py4CastLoss):
    """Per-feature weighted MSE / L1, mean over interior pixels (reference losses.py:103-169)."""

    def prepare(self, lm, interior_mask, dataset_info) -> None:
        exponent = 2.0 if self.loss_name == "MSELoss" else 1.0
        loss_state_weight = {}
        for name in dataset_info.state_weights:
            loss_state_weight[name] = dataset_info.state_weights[name] / (
                dataset_info.diff_stats[name]["std"] ** exponent
            )
        self.register_loss_state_buffers(lm, interior_mask, loss_state_weight, squeeze_mask=True)
        self.lm = lm

    def forward(self, prediction: NamedTensor, target: NamedTensor, mask: Optional[torch.Tensor] = None,
                reduce_spatial_dim: bool = True) -> torch.Tensor:
        """(B,T,N,F) or (B,T,H,W,F) -> (B,T); (B,T,*spatial) when reduce_spatial_dim=False."""
        w = self.weights(tuple(prediction.feature_names), prediction.device)
        form = FORM_WEIGHTED if reduce_spatial_dim else FORM_MAP
        return _LossFn.apply(
            prediction.tensor, target.tensor, mask, self._interior_flat(prediction.device), w,
            self.num_interior, self.kind, form,
        )


class ScaledLoss(Py4CastLoss):
    """std-scaled RMSE / MAE per feature (reference losses.py:172-210) -> (B,T,F)."""

    def prepare(self, lm, interior_mask, dataset_info) -> None:
        loss_state_weight = {}
        for name in dataset_info.state_weights:
            loss_state_weight[name] = dataset_info.stats[name]["std"]
        self.register_loss_state_buffers(lm, interior_mask, loss_state_weight)
        self.lm = lm

    def forward(self, prediction: NamedTensor, target: NamedTensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        w = self.weights(tuple(prediction.feature_names), prediction.device)
        return _LossFn.apply(
            prediction.tensor, target.tensor, mask, self._interior_flat(prediction.device), w,
            self.num_interior, self.kind, FORM_SCALED,
        )


class CombinedLoss(Py4CastLoss):
    """Weighted sum of Py4CastLoss terms (reference losses.py:263-307)."""

    def __init__(self, losses_config: list):
        self.losses = []
        for loss_conf in losses_config:
            kls = loss_conf["class"]
            LossClass = globals()[kls] if isinstance(kls, str) else kls
            weight = loss_conf.get("weight", 1.0)
            kwargs = loss_conf.get("params", {})
            self.losses.append((LossClass(**kwargs), weight))

    def prepare(self, lm, interior_mask, dataset_info):
        for loss, _ in self.losses:
            if hasattr(loss, "prepare"):
                loss.prepare(lm, interior_mask, dataset_info)

    def forward(self, prediction: NamedTensor, target: NamedTensor, mask: Optional[torch.Tensor] = None,
                **kwargs) -> torch.Tensor:
        total = None
        for loss, weight in self.losses:
            term = loss(prediction, target, mask, **kwargs)
            term = term if weight == 1.0 else weight * term
            total = term if total is None else total + term
        return total
