"""
py4cast_b200 — B200 (sm_100a) implementation of py4cast's autoregressive-training hot path:
GraphLAM / HiLAM message passing, the scaled / differential AR state update and the
Weighted / Scaled losses, behind the reference's own plugin surface (model registry + settings
dataclasses, nn.Module forward over NamedTensor batches, Py4CastLoss.prepare/forward,
common_step(batch, batch_idx, phase)).  See DESIGN.md and INTEGRATION.md.

Importing the package never touches CUDA; the kernels (py4cast_b200/lib/libpy4cast_b200.so,
C ABI in include/py4cast_b200.h) are loaded on first use and there is no CPU fallback.
"""

from .namedtensor import NamedTensor  # noqa: F401

__version__ = "0.1.0"
