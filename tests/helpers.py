"""Shared helpers of the test-suite (oracle access lives HERE, never in the package)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FP32_TOL = 1e-5  # north_star: 1e-5 relative in fp32
BF16_TOL = 2e-2  # north_star: 2e-2 in bf16


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b|| / ||b||  (Frobenius), b = reference."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    den = b.norm().item()
    if den == 0.0:
        return (a - b).norm().item()
    return ((a - b).norm() / den).item()


REPORT = []  # (test id, what, error, tolerance) — dumped by conftest at session end


def assert_close(a, b, tol, what=""):
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    e = rel_err(a, b)
    REPORT.append((os.environ.get("PYTEST_CURRENT_TEST", "?").split(" ")[0], what, e, tol))
    assert e <= tol, f"{what}: relative error {e:.3e} > {tol:.1e}"
    return e


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)


def pixel_meshgrid(H, W):
    """tests/test_models.py:54-60 (FakeStatics.meshgrid) of the reference."""
    xx, yy = np.meshgrid(np.arange(0, W, 1), np.arange(0, H, 1))
    return torch.from_numpy(np.asarray([xx, yy]))
