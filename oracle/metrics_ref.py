"""
TEST INFRASTRUCTURE ONLY (oracle): CPU restatement of the reference's validation metrics
(py4cast/metrics.py) — the numerics only, no torchmetrics state, no figures:

  * power_spectral_density / dct_var / radial_bin_dct   metrics.py:253-352  (used by MetricPSDK.add_psd
    :121-144 and MetricPSDVar.update :204-213)
  * psd_log_rmse                                         metrics.py:199-213  (MetricPSDVar)
  * acc                                                  metrics.py:407-420  (MetricACC.update)

Pinned against the verbatim reference by tests/golden/metrics.pt (oracle/make_golden_metrics.py loads
/root/reference/py4cast/metrics.py unmodified under import stubs for torchmetrics / plots).  Prepared in
round 1 for SURVEY 8(f) rank 3; the CUDA side is not written yet.  Only tests/ may import this module.

Faithful to a quirk of the reference: radial_bin_dct indexes the FLATTENED variance spectrum with
2 r (metrics.py:311-315), so the value binned at radius r depends on r only:
    profile[r] = sig.ravel()[2 r] + 0.5 sig.ravel()[2 r - 1] + 0.5 sig.ravel()[2 r + 1],   r < Rmax
(index -1 wraps to the last element for r = 0), i.e. only the first 2 Rmax + 1 flat entries of the
2-D DCT spectrum (its first row for the usual shapes) and the very last one are ever read.
"""
import numpy as np
import torch
from scipy.fftpack import dct


def dct_2d(x: np.ndarray) -> np.ndarray:
    """metrics.py:253-268: orthonormal DCT-II along the last two axes of (N, N) or (B, N, N)."""
    if x.ndim == 3:
        return dct(dct(x.transpose((0, 2, 1)), norm="ortho").transpose((0, 2, 1)), norm="ortho")
    return dct(dct(x.T, norm="ortho").T, norm="ortho")


def variance_spectrum(x: np.ndarray) -> np.ndarray:
    """metrics.py:271-286 (`dct_var`): (1 / n^2) DCT2(x)^2 with n = last dimension."""
    return (1.0 / x.shape[-1] ** 2) * dct_2d(x) ** 2


def radial_profile(sig: np.ndarray) -> np.ndarray:
    """metrics.py:289-321 with center = (shape[0] // 2, shape[1] // 2) (:345)."""
    cy, cx = sig.shape[0] // 2, sig.shape[1] // 2
    y, x = np.indices(sig.shape)
    r = np.sqrt((x - cy) ** 2 + (y - cx) ** 2).astype(int)  # center[0] pairs with x, center[1] with y (:303)
    rmax = min(x.max(), y.max(), r.max()) // 2
    flat = sig.ravel()
    present = np.bincount(r.ravel()[r.ravel() < rmax], minlength=0)
    out = np.empty(len(present), dtype=sig.dtype)
    for k in range(len(present)):
        out[k] = flat[2 * k] + 0.5 * flat[2 * k - 1] + 0.5 * flat[2 * k + 1] if present[k] else np.nan
    return out


def power_spectral_density(x: np.ndarray) -> np.ndarray:
    """metrics.py:324-352: x (B, C, N, M) -> (C, Rmax)."""
    return np.stack([radial_profile(variance_spectrum(x[:, c]).mean(axis=0)) for c in range(x.shape[1])], axis=0)


def psd_log_rmse(pred: torch.Tensor, target: torch.Tensor, mask: torch.Tensor, pred_step: int) -> np.ndarray:
    """metrics.py:186-213: per feature, RMSE between log10 PSDs of (pred, target) * mask at one step.
    pred / target / mask: (B, T, H, W, F)."""
    perm = lambda t: t.permute(0, -1, 2, 3, 1)  # noqa: E731  (B, F, H, W, T)
    yp, yt = perm(pred) * perm(mask), perm(target) * perm(mask)
    res = np.zeros((yp.shape[1],))
    for c in range(yp.shape[1]):
        pt = power_spectral_density(yt[:, c:c + 1, :, :, pred_step].numpy())
        pp = power_spectral_density(yp[:, c:c + 1, :, :, pred_step].numpy())
        res[c] = np.sqrt(np.mean((np.log10(pt) - np.log10(pp)) ** 2))
    return res


def acc(pred: torch.Tensor, target: torch.Tensor, mask: torch.Tensor, climate_means: torch.Tensor,
        spatial_dims) -> torch.Tensor:
    """metrics.py:407-420: anomaly correlation coefficient, (T, F) after the batch mean."""
    num = ((pred - climate_means) * (target - climate_means) * mask).mean(dim=spatial_dims)
    den = (((pred - climate_means) * mask) ** 2).mean(dim=spatial_dims) * (
        ((target - climate_means) * mask) ** 2
    ).mean(dim=spatial_dims)
    return torch.mean(num / torch.sqrt(den), dim=0)
