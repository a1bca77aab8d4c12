"""
Dataset statistics on the device: the numbers of reference
py4cast/datasets/compute_dataset_stats.py — `parameters_stats.pt` (mean / std / min / max per feature)
and `diff_stats.pt` (mean / std of the one-step time differences, the `diff_std` the losses and the AR
update divide / multiply by) — from an iterable of batches, one pass per batch through
`p4c_feature_stats` (NaN-ignoring mean, mean of squares, min, max per (sample, feature), fp64
accumulation) instead of ~10 full-size temporaries per batch.  Same function names, the same
accumulation over batches and the same quirks as the reference:

  * `sum_means += nansum(nanmean over space)` per sample, divided by the SAMPLE count at the end
    (:47-48, :62-64): every sample weighs the same whatever its NaN count;
  * the running min / max only ever looks at sample 0 of each batch (`.values[0]`, :50-55), after an
    initial min / max over all samples of the first batch (:24-25).

`unnormalize_` is the in-place un-normalisation of a prediction before it is written out
(lightning.py:1162-1169).  No CPU fallback: CUDA tensors only.
"""

from __future__ import annotations

from typing import Dict, Iterable, Literal, Optional

import torch

from . import _lib
from ._lib import check, ptr, stream
from .namedtensor import NamedTensor


def feature_stats(x: torch.Tensor, prev: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x (B, ..., F) fp32 CUDA [minus prev] -> (B, F, 4): nanmean, nanmean of squares, min, max over the
    flattened middle dimensions."""
    if not x.is_cuda:
        raise _lib.P4CError("py4cast_b200.dataset_stats: CUDA tensors required (no CPU fallback)")
    B, F = x.shape[0], x.shape[-1]
    xf = x.to(torch.float32).contiguous()
    pf = None
    if prev is not None:
        if prev.shape != x.shape:
            raise ValueError("feature_stats: x and prev must have the same shape")
        pf = prev.to(torch.float32).contiguous()
    X = xf.numel() // (B * F)
    lib = _lib.load()
    ws = torch.empty(max(1, lib.p4c_feature_stats_partial_bytes(B, X, F)), dtype=torch.uint8, device=x.device)
    out = torch.empty(B, F, 4, dtype=torch.float32, device=x.device)
    check(lib.p4c_feature_stats(ptr(xf), ptr(pf), B, X, F, ptr(ws), ptr(out), stream()), "p4c_feature_stats")
    return out


def compute_mean_std_min_max(batches: Iterable, type_tensor: Literal["inputs", "outputs", "forcing"]) -> Dict[str, dict]:
    """compute_dataset_stats.py:11-75 over an iterable of (un-standardised) ItemBatch on the device."""
    sum_means = sum_squares = best_min = best_max = None
    counter, names = 0, None
    for batch in batches:
        nt: NamedTensor = getattr(batch, type_tensor)
        st = feature_stats(nt.tensor)  # (B, F, 4)
        if names is None:
            names = list(nt.feature_names)
            best_min, best_max = st[:, :, 2].min(dim=0).values, st[:, :, 3].max(dim=0).values  # :24-25
        counter += st.shape[0]
        m, q = torch.nansum(st[:, :, 0], dim=0), torch.nansum(st[:, :, 1], dim=0)
        sum_means = m if sum_means is None else sum_means + m
        sum_squares = q if sum_squares is None else sum_squares + q
        best_min = torch.minimum(best_min, st[0, :, 2])  # sample 0 only, as the reference (:50-51)
        best_max = torch.maximum(best_max, st[0, :, 3])
    if names is None:
        raise ValueError("compute_mean_std_min_max: no batch")
    mean = sum_means / counter
    std = torch.sqrt(sum_squares / counter - mean**2)
    return {n: {"mean": mean[i], "std": std[i], "min": best_min[i], "max": best_max[i]} for i, n in enumerate(names)}


def compute_parameters_stats(batches_fn) -> Dict[str, dict]:
    """compute_dataset_stats.py:78-92: `batches_fn()` returns a fresh iterable of batches; the first
    occurrence of a feature wins.  Returns the dict the reference saves as parameters_stats.pt."""
    all_stats: Dict[str, dict] = {}
    for type_tensor in ("inputs", "outputs", "forcing"):
        for feature, stats in compute_mean_std_min_max(batches_fn(), type_tensor).items():
            all_stats.setdefault(feature, stats)
    return all_stats


def compute_time_step_stats(batches: Iterable) -> Dict[str, dict]:
    """compute_dataset_stats.py:95-127 over (standardised) batches: mean / std of in_out[t + 1] - in_out[t];
    forcing features get the fixed mean 0 / std 1.  Returns the dict saved as diff_stats.pt."""
    sum_means = sum_squares = None
    counter, names, fnames = 0, None, []
    for batch in batches:
        in_out = torch.cat([batch.inputs.tensor, batch.outputs.tensor], dim=1)
        st = feature_stats(in_out[:, 1:], in_out[:, :-1])
        names, fnames = list(batch.inputs.feature_names), list(batch.forcing.feature_names)
        counter += st.shape[0]
        m, q = torch.nansum(st[:, :, 0], dim=0), torch.nansum(st[:, :, 1], dim=0)
        sum_means = m if sum_means is None else sum_means + m
        sum_squares = q if sum_squares is None else sum_squares + q
    if names is None:
        raise ValueError("compute_time_step_stats: no batch")
    mean = sum_means / counter
    std = torch.sqrt(sum_squares / counter - mean**2)
    out = {n: {"mean": mean[i], "std": std[i]} for i, n in enumerate(names)}
    for n in fnames:
        out[n] = {"mean": torch.tensor(0), "std": torch.tensor(1)}
    return out


def unnormalize_(preds: NamedTensor, stats) -> NamedTensor:
    """preds[..., f] = preds[..., f] * std_f + mean_f in place (lightning.py:1162-1169), one kernel."""
    t = preds.tensor
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise _lib.P4CError("unnormalize_: contiguous fp32 CUDA tensor required (no CPU fallback)")
    names = list(preds.feature_names)
    std = stats.to_list("std", names).to(t.device, torch.float32).contiguous()
    mean = stats.to_list("mean", names).to(t.device, torch.float32).contiguous()
    F = t.shape[-1]
    check(_lib.load().p4c_scale_shift(ptr(t), t.numel() // F, F, ptr(std), ptr(mean), stream()), "p4c_scale_shift")
    return preds
