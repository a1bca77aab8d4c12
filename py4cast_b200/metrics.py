"""
Validation metrics on the device.

`MetricACC` mirrors reference py4cast/metrics.py:355-455 (same constructor argument, `update(preds,
target, mask)`, `compute(prefix)`, `reset()`; no torchmetrics dependency: the two states it keeps,
`sum_acc` and `step_count`, are plain tensors and `compute` returns the same dict of per-feature,
per-step scalars).  The three spatial reductions of `update` run in one pass over prediction and
target (`p4c_acc_forward`) instead of ~10 full-size temporaries; parity with the verbatim reference
at 1e-5 on golden vectors (tests/test_metrics_oracle.py).  PSD metrics (metrics.py:13-352): oracle only
so far (`oracle/metrics_ref.py`).
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
