"""
Validation metrics on the device.

`MetricACC` mirrors reference py4cast/metrics.py:355-455 (same constructor argument, `update(preds,
target, mask)`, `compute(prefix)`, `reset()`; no torchmetrics dependency: the two states it keeps,
`sum_acc` and `step_count`, are plain tensors and `compute` returns the same dict of per-feature,
per-step scalars).  The three spatial reductions of `update` run in one pass over prediction and
target (`p4c_acc_forward`) instead of ~10 full-size temporaries; parity with the verbatim reference
at 1e-5 on golden vectors (tests/test_metrics_oracle.py).

`power_spectral_density`, `MetricPSDVar` and `MetricPSDK` mirror metrics.py:13-352: the reference moves
every (sample, feature) field to the host, takes a scipy 2-D DCT and bins it; because its binning reads
the flattened spectrum at 2 r - 1, 2 r, 2 r + 1 only (metrics.py:311-315), the same numbers come from ONE
pass over the fields on the device (`p4c_psd_first_row`: column sums + a DCT of the sums), all features and
both tensors at once, with no host round trip.  Parity with the verbatim reference on golden vectors
(tests/test_gpu_metrics.py).  The figure of MetricPSDK.compute (matplotlib) is out of scope: `compute`
returns the mean spectra.
"""

from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from ._lib import check, ptr, stream
from .namedtensor import NamedTensor


def acc_per_sample(pred: torch.Tensor, target: torch.Tensor, mask: Optional[torch.Tensor],
                   climate_means: torch.Tensor) -> torch.Tensor:
    """(B, T, *space, F) fp32 CUDA tensors -> (B, T, F) anomaly correlation coefficients."""
    if not (pred.is_cuda and target.is_cuda):
        raise _lib.P4CError("py4cast_b200.metrics: CUDA tensors required (no CPU fallback)")
    if pred.shape != target.shape:
        raise ValueError("preds and target must have the same shape")
    B, T, F = pred.shape[0], pred.shape[1], pred.shape[-1]
    nodes = pred.numel() // (B * T * F)
    p, t = pred.float().contiguous(), target.float().contiguous()
    m = None if mask is None else mask.to(torch.float32).expand_as(p).contiguous()
    c = climate_means.to(p.device, torch.float32).contiguous()
    lib = _lib.load()
    ws = torch.empty(max(1, lib.p4c_acc_partial_floats(B * T, nodes, F)), dtype=torch.float32, device=p.device)
    out = torch.empty(B, T, F, dtype=torch.float32, device=p.device)
    check(lib.p4c_acc_forward(ptr(p), ptr(t), ptr(m), ptr(c), B * T, nodes, F, ptr(ws), ptr(out), stream()),
          "p4c_acc_forward")
    return out


class MetricACC:
    """Anomaly correlation coefficient against per-feature climate means (one scalar per field, as in
    the reference), averaged over the batch at every `update` and over the updates at `compute`."""

    def __init__(self, dataset_info):
        names = dataset_info.shortnames["input_output"] + dataset_info.shortnames["output"]
        self.climate_means = dataset_info.stats.to_list("mean", names)
        self.reset()

    def reset(self) -> None:
        self.sum_acc: Optional[torch.Tensor] = None
        self.step_count = 0

    def update(self, preds: NamedTensor, target: NamedTensor, mask: Optional[torch.Tensor], *args) -> None:
        if self.step_count == 0:
            self.feature_names = preds.feature_names
            self.pred_steps = preds.tensor.shape[1]
        res = acc_per_sample(preds.tensor, target.tensor, mask, self.climate_means).mean(dim=0)  # (T, F)
        self.sum_acc = res if self.sum_acc is None else self.sum_acc + res
        self.step_count += 1

    def compute(self, prefix: str = "val") -> dict:
        mean_acc = self.sum_acc / self.step_count
        out = {
            f"{prefix}_acc/{name}_step{j}": mean_acc[j, i]
            for i, name in enumerate(self.feature_names)
            for j in range(self.pred_steps)
        }
        self.reset()
        return out


# ------------------------------------------------------------------------------------- PSD
def _psd_bins(H: int, W: int):
    """Radial bins of reference metrics.py:289-321 for an (H, W) spectrum, as a bool array `present`
    (bin r holds at least one pixel; empty bins are 0 / 0 = NaN in the reference): radii are taken around
    (cy, cx) = (H // 2, W // 2) with the axes swapped (:303) and kept below min(W - 1, H - 1, r.max()) // 2."""
    import numpy as np

    cy, cx = H // 2, W // 2
    y, x = np.indices((H, W))
    r = np.sqrt((x - cy) ** 2 + (y - cx) ** 2).astype(int)
    rmax = int(min(x.max(), y.max(), r.max()) // 2)
    return np.bincount(r.ravel()[r.ravel() < rmax], minlength=0) > 0


_PSD_CONST: dict = {}


def power_spectral_density(x: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x (B, H, W, F) fp32 CUDA (features last, one prediction step) -> (F, R): the radial PSD profile of
    reference `power_spectral_density` (metrics.py:324-352) of every feature, batch-averaged (NaN in the
    bins the reference leaves empty)."""
    if not x.is_cuda:
        raise _lib.P4CError("py4cast_b200.metrics: CUDA tensors required (no CPU fallback)")
    B, H, W, F = x.shape
    key = (H, W, str(x.device))
    if key not in _PSD_CONST:
        import math

        present = torch.from_numpy(_psd_bins(H, W)).to(torch.uint8).to(x.device)  # one byte per bin
        wy = torch.tensor([math.cos(math.pi * (2 * y + 1) * (H - 1) / (2 * H)) for y in range(H)],
                          dtype=torch.float32, device=x.device)
        R = present.numel()
        J = 2 * R  # flat spectrum entries 0 .. 2 R - 1 (+ the last one)
        if J > W or R < 1:
            raise NotImplementedError("PSD: the reference's flat indexing leaves the first spectrum row for this shape")
        xs = torch.arange(W, dtype=torch.float64).unsqueeze(1)
        ks = torch.cat([torch.arange(J, dtype=torch.float64), torch.tensor([W - 1.0], dtype=torch.float64)]).unsqueeze(0)
        ctab = torch.cos(math.pi * (2 * xs + 1) * ks / (2 * W)).to(torch.float32).contiguous().to(x.device)
        _PSD_CONST[key] = (present, wy, ctab)
    present, wy, ctab = _PSD_CONST[key]
    R = present.numel()
    J = 2 * R
    xf = x.float().contiguous()
    m = None if mask is None else mask.to(torch.float32).expand_as(xf).contiguous()
    lib = _lib.load()
    ws = torch.empty(max(1, lib.p4c_psd_partial_floats(B, H, W, F)), dtype=torch.float32, device=x.device)
    spec = torch.empty(B, F, J + 1, dtype=torch.float32, device=x.device)
    check(lib.p4c_psd_first_row(ptr(xf), ptr(m), ptr(wy), ptr(ctab), B, H, W, F, J, ptr(ws), ptr(spec), stream()),
          "p4c_psd_first_row")
    # batch mean of the variance spectrum (metrics.py:343) and the radial bins (center + half of both neighbours,
    # index -1 wrapping to the last entry; NaN in the bins the reference leaves empty): one small kernel
    prof = torch.empty(F, R, dtype=torch.float32, device=x.device)
    check(lib.p4c_psd_profile(ptr(spec), ptr(present), B, F, R, ptr(prof), stream()), "p4c_psd_profile")
    return prof


def _unflatten_for_psd(nt: NamedTensor, shape) -> torch.Tensor:
    t = nt.tensor
    if shape is not None and t.dim() == 4:  # (B, T, ngrid, F) -> (B, T, H, W, F), metrics.py:52-60
        t = t.unflatten(2, tuple(shape[2:4]))
    return t


class MetricPSDVar:
    """RMSE between the log10 PSDs of prediction and target, per feature (reference metrics.py:147-250):
    same constructor argument, `update(preds, targets, mask, shape)`, `compute(prefix)`, `reset()`."""

    def __init__(self, pred_step: int = 0):
        self.pred_step = pred_step
        self.reset()

    def reset(self) -> None:
        self.sum_rmse: Optional[torch.Tensor] = None
        self.step_count = 0

    def update(self, preds: NamedTensor, targets: NamedTensor, mask: Optional[torch.Tensor], shape=None) -> None:
        if self.step_count == 0:
            self.feature_names = preds.feature_names
        p, t = _unflatten_for_psd(preds, shape), _unflatten_for_psd(targets, shape)
        if p.shape != t.shape:
            raise ValueError("y_pred and y_true must have the same shape")
        m = None
        if mask is not None:
            m = mask.unflatten(2, tuple(shape[2:4])) if (shape is not None and mask.dim() == 4) else mask
            m = m[:, self.pred_step]
        psd_t = power_spectral_density(t[:, self.pred_step], m)
        psd_p = power_spectral_density(p[:, self.pred_step], m)
        res = torch.sqrt(torch.mean((torch.log10(psd_t) - torch.log10(psd_p)) ** 2, dim=1))  # (F,)
        self.sum_rmse = res if self.sum_rmse is None else self.sum_rmse + res
        self.step_count += 1

    def compute(self, prefix: str = "val") -> dict:
        mean_psd_rmse = self.sum_rmse / self.step_count
        out = {f"{prefix}_rmse_psd/{name}": mean_psd_rmse[i] for i, name in enumerate(self.feature_names)}
        self.reset()
        return out


class MetricPSDK:
    """Mean PSD of prediction and target as a function of the wave number, per feature (reference
    metrics.py:13-144).  `compute` returns {feature: (psd_pred, psd_target)} (each (Rmax,)); the matplotlib
    figure the reference builds from them (:96-119) is presentation and stays with the caller."""

    def __init__(self, save_path=None, pred_step: int = 0):
        self.save_path, self.pred_step = save_path, pred_step
        self.reset()

    def reset(self) -> None:
        self.sum_psd_pred: Optional[torch.Tensor] = None
        self.sum_psd_target: Optional[torch.Tensor] = None
        self.step_count = 0

    def update(self, preds: NamedTensor, targets: NamedTensor, mask: Optional[torch.Tensor], shape=None) -> None:
        if preds.tensor.shape != targets.tensor.shape:
            raise ValueError("preds and targets must have the same shape")
        if self.step_count == 0:
            self.feature_names = preds.feature_names
        p, t = _unflatten_for_psd(preds, shape), _unflatten_for_psd(targets, shape)
        m = None
        if mask is not None:
            m = mask.unflatten(2, tuple(shape[2:4])) if (shape is not None and mask.dim() == 4) else mask
            m = m[:, self.pred_step]
        pp, pt = power_spectral_density(p[:, self.pred_step], m), power_spectral_density(t[:, self.pred_step], m)
        self.sum_psd_pred = pp if self.sum_psd_pred is None else self.sum_psd_pred + pp
        self.sum_psd_target = pt if self.sum_psd_target is None else self.sum_psd_target + pt
        self.step_count += 1

    def compute(self, prefix: str = "val") -> dict:
        out = {
            name: (self.sum_psd_pred[i] / self.step_count, self.sum_psd_target[i] / self.step_count)
            for i, name in enumerate(self.feature_names)
        }
        self.reset()
        return out
