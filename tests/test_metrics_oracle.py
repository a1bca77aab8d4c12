"""Validation metrics: the oracle restatement of the reference's PSD / ACC numerics
against golden vectors from the VERBATIM reference metrics.py (oracle/make_golden_metrics.py), and the
ACC reduction kernel against the same vectors.  PSD has no CUDA side yet."""
import numpy as np
import pytest
import torch

from helpers import assert_close, load_golden


@pytest.mark.parametrize("name", ["square", "wide", "tall"])
def test_psd_oracle_matches_reference_golden(name):
    from oracle import metrics_ref

    c = load_golden("metrics.pt")["psd"][name]
    got = metrics_ref.power_spectral_density(c["x"].numpy())
    exp = c["psd"].numpy()
    assert got.shape == exp.shape
    assert np.array_equal(np.isnan(got), np.isnan(exp))  # radii without pixels are 0 / 0 in the reference too
    ok = ~np.isnan(exp)
    assert np.allclose(got[ok], exp[ok], rtol=1e-12, atol=0)


def test_psd_reads_only_the_head_of_the_flat_spectrum():
    """The quirk the GPU kernel will have to reproduce: the profile is a function of the first flat
    entries of the variance spectrum (and the last one), not of a true radial average."""
    from oracle import metrics_ref

    g = np.random.default_rng(0)
    sig = g.random((24, 40))
    base = metrics_ref.radial_profile(sig)
    n = len(base)
    other = sig.copy()
    other.ravel()[2 * n + 2:-1] = g.random(other.size - 2 * n - 3)  # everything but the head and the last entry
    assert np.array_equal(metrics_ref.radial_profile(other), base)


@pytest.mark.parametrize("name", ["graph", "grid"])
def test_acc_oracle_matches_reference_golden(name):
    from oracle import metrics_ref

    c = load_golden("metrics.pt")["acc"][name]
    got = metrics_ref.acc(c["pred"], c["target"], c["mask"], c["climate_means"], tuple(c["spatial_dims"]))
    assert got.shape == c["acc"].shape
    assert_close(got, c["acc"], 1e-6, f"acc {name}")
    assert float(got.min()) > 0.5  # prediction = target + noise: strongly correlated anomalies


def test_psd_log_rmse_is_zero_for_identical_fields():
    from oracle import metrics_ref

    x = torch.randn(2, 2, 16, 16, 3, dtype=torch.float64)
    m = torch.ones_like(x)
    assert np.allclose(metrics_ref.psd_log_rmse(x, x, m, 1), 0.0)
    assert (metrics_ref.psd_log_rmse(x, 2 * x, m, 0) > 0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["graph", "grid"])
def test_gpu_acc_vs_reference_golden(cuda, name):
    from py4cast_b200.metrics import acc_per_sample

    c = load_golden("metrics.pt")["acc"][name]
    per_sample = acc_per_sample(c["pred"].to(cuda), c["target"].to(cuda), c["mask"].to(cuda), c["climate_means"])
    assert_close(per_sample.mean(dim=0).cpu(), c["acc"], 1e-5, f"acc {name}")


def test_metric_acc_refuses_cpu_tensors():
    from py4cast_b200 import _lib
    from py4cast_b200.metrics import acc_per_sample

    x = torch.randn(1, 1, 8, 2)
    with pytest.raises(_lib.P4CError):
        acc_per_sample(x, x, None, torch.zeros(2))
