"""
ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under py4cast_b200/ may import this.

Loads the reference's OWN, UNMODIFIED hot-path sources

    /root/reference/py4cast/losses.py      (Py4CastLoss, WeightedLoss, ScaledLoss, CombinedLoss)
    /root/reference/py4cast/lightning.py   (AutoRegressiveLightning.common_step/_common_step/
                                            _next_x/_step_diffs/_strategy_params/get_mask_on_nan)

by file path, after injecting stub `sys.modules` entries for their peripheral
imports (lightning, mlflow, matplotlib, torchinfo, mfai, sibling py4cast
modules) that are not installed in this image.  No reference
source is copied: the files are exec'd where they lie.

This only works where /root/reference exists (the build container).  It is
used by oracle/make_golden.py to generate tests/golden/*.pt and by the
`not gpu` tests that pin oracle/ar_ref.py against the real reference.
"""

from __future__ import annotations

import contextlib
import importlib.util
import os
import sys
import types
from dataclasses import dataclass
from functools import cached_property
from typing import Any

import torch

from . import mfai_shim

REFERENCE_ROOT = os.environ.get("PY4CAST_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "py4cast", "lightning.py"))


# --------------------------------------------------------------------------
# Stand-in for py4cast.datasets.base.{Item, ItemBatch} (reference base.py:40-170):
# same three NamedTensor fields and the same three cached properties.  The real
# file cannot be exec'd here (imports gif, cartopy, matplotlib.pyplot ...).
# --------------------------------------------------------------------------
@dataclass
class ItemBatch:
    inputs: Any
    forcing: Any
    outputs: Any
    validity_times: Any = None

    @cached_property
    def batch_size(self):
        return self.outputs.dim_size("batch")

    @cached_property
    def num_input_steps(self):
        if self.inputs.dim_size("timestep") is None:
            return self.outputs.dim_size("timestep")
        return self.inputs.dim_size("timestep")

    @cached_property
    def num_pred_steps(self):
        return self.outputs.dim_size("timestep")


class DictStats:
    """In-memory equivalent of py4cast.datasets.access.Stats (access.py:355-390)."""

    def __init__(self, stats: dict):
        self.stats = stats

    def items(self):
        return self.stats.items()

    def __getitem__(self, name):
        return self.stats[name]

    def to_list(self, stat_name, shortnames, dtype=torch.float32):
        if len(shortnames) > 0:
            return torch.stack([self[n][stat_name] for n in shortnames], dim=0).type(dtype)
        return []


def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def _stub_modules(named_tensor_cls) -> dict:
    class LightningModule(torch.nn.Module):
        pass

    class LightningDataModule:
        def __init__(self, *a, **k):
            pass

    class _Anything:
        def __init__(self, *a, **k):
            pass

    def rank_zero_only(fn):
        return fn

    str_to_dtype = {
        "bf16-true": torch.bfloat16,
        "16-true": torch.float16,
        "32-true": torch.float32,
        "64-true": torch.float64,
    }

    pl = _mod("lightning.pytorch", LightningModule=LightningModule)
    mods = {
        "lightning": _mod(
            "lightning",
            LightningModule=LightningModule,
            LightningDataModule=LightningDataModule,
            pytorch=pl,
        ),
        "lightning.pytorch": pl,
        "lightning.pytorch.loggers": _mod("lightning.pytorch.loggers", MLFlowLogger=_Anything),
        "lightning.pytorch.utilities": _mod(
            "lightning.pytorch.utilities", rank_zero_only=rank_zero_only
        ),
        "matplotlib": _mod("matplotlib"),
        "mlflow": _mod("mlflow"),
        "mlflow.pytorch": _mod("mlflow.pytorch"),
        "mlflow.models": _mod("mlflow.models"),
        "mlflow.models.signature": _mod("mlflow.models.signature", infer_signature=None),
        "torchinfo": _mod("torchinfo", summary=lambda *a, **k: None),
        "mfai": _mod("mfai"),
        "mfai.pytorch": _mod("mfai.pytorch"),
        "mfai.pytorch.namedtensor": _mod(
            "mfai.pytorch.namedtensor", NamedTensor=named_tensor_cls
        ),
        "mfai.pytorch.losses": _mod("mfai.pytorch.losses"),
        "mfai.pytorch.losses.perceptual": _mod(
            "mfai.pytorch.losses.perceptual", PerceptualLoss=object
        ),
        "mfai.pytorch.models": _mod("mfai.pytorch.models", registry={}),
        "mfai.pytorch.models.base": _mod(
            "mfai.pytorch.models.base",
            ModelType=mfai_shim.ModelType,
            ModelABC=mfai_shim.ModelABC,
        ),
        "mfai.pytorch.models.utils": _mod(
            "mfai.pytorch.models.utils",
            expand_to_batch=mfai_shim.expand_to_batch,
            features_last_to_second=mfai_shim.features_last_to_second,
            features_second_to_last=mfai_shim.features_second_to_last,
        ),
        "py4cast": _mod("py4cast"),
        "py4cast.datasets": _mod("py4cast.datasets", get_datasets=None),
        "py4cast.datasets.base": _mod(
            "py4cast.datasets.base",
            DatasetInfo=object,
            ItemBatch=ItemBatch,
            NamedTensor=named_tensor_cls,
            Statics=object,
        ),
        "py4cast.io": _mod("py4cast.io"),
        "py4cast.io.outputs": _mod(
            "py4cast.io.outputs",
            OutputSavingSettings=object,
            save_gifs=None,
            save_named_tensors_to_grib=None,
        ),
        "py4cast.metrics": _mod(
            "py4cast.metrics", MetricACC=None, MetricPSDK=None, MetricPSDVar=None
        ),
        "py4cast.models": _mod(
            "py4cast.models",
            build_model_from_settings=None,
            get_model_kls_and_settings=None,
            registry={"HalfUNet": None},
        ),
        "py4cast.plots": _mod(
            "py4cast.plots",
            PredictionEpochPlot=None,
            PredictionTimestepPlot=None,
            SpatialErrorPlot=None,
            StateErrorPlot=None,
        ),
        "py4cast.utils": _mod("py4cast.utils", str_to_dtype=str_to_dtype),
    }
    mods["py4cast"].__path__ = []  # mark as package
    mods["lightning"].__path__ = []
    mods["mlflow"].__path__ = []
    mods["mfai"].__path__ = []
    return mods


@contextlib.contextmanager
def _patched_sys_modules(mods: dict):
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        yield
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


_CACHE: dict = {}


def load_reference(named_tensor_cls=None):
    """
    Returns (ref_losses_module, ref_lightning_module): the reference files
    executed verbatim.  `named_tensor_cls` lets a test run the reference on top
    of the product's NamedTensor instead of the oracle's own shim.
    """
    named_tensor_cls = named_tensor_cls or mfai_shim.NamedTensor
    key = id(named_tensor_cls)
    if key in _CACHE:
        return _CACHE[key]
    if not reference_available():
        raise FileNotFoundError(
            f"reference sources not found under {REFERENCE_ROOT}; the verbatim "
            "reference can only be executed in the build container"
        )
    mods = _stub_modules(named_tensor_cls)
    with _patched_sys_modules(mods):
        out = []
        for name in ("losses", "lightning"):
            path = os.path.join(REFERENCE_ROOT, "py4cast", f"{name}.py")
            spec = importlib.util.spec_from_file_location(f"py4cast.{name}", path)
            module = importlib.util.module_from_spec(spec)
            sys.modules[f"py4cast.{name}"] = module
            try:
                spec.loader.exec_module(module)
            finally:
                if name == "lightning":
                    sys.modules.pop("py4cast.losses", None)
                    sys.modules.pop("py4cast.lightning", None)
            out.append(module)
    _CACHE[key] = tuple(out)
    return _CACHE[key]


class _Trainer:
    def __init__(self, precision: str):
        self.precision = precision


def make_reference_module(
    model: torch.nn.Module,
    *,
    training_strategy: str,
    border_mask: torch.Tensor,
    grid_static_features: torch.Tensor,
    diff_stats: DictStats,
    stats: DictStats | None = None,
    num_inter_steps: int = 1,
    precision: str = "32-true",
    named_tensor_cls=None,
    mask_on_nan: bool = False,
    mask_ratio: float = 0,
):
    """
    Builds a reference `AutoRegressiveLightning` WITHOUT running its __init__
    (which needs a dataset, the CLI and Lightning): allocates it with __new__,
    initialises nn.Module, and sets exactly the attributes / buffers
    `_common_step` reads (lightning.py:285-300, 519-676).

    border_mask: (H,W,1) for conv models or (N,1) for graph models, float {0,1}.
    grid_static_features: (B, *spatial, F_s) — already expanded to batch
    (lightning.py:296-300).
    """
    _, ref_lightning = load_reference(named_tensor_cls)
    lm = ref_lightning.AutoRegressiveLightning.__new__(ref_lightning.AutoRegressiveLightning)
    torch.nn.Module.__init__(lm)
    lm.model = model
    lm.training_strategy = training_strategy
    lm.num_inter_steps = num_inter_steps
    lm.mask_on_nan = mask_on_nan
    lm.channels_last = False
    lm.mask_ratio = mask_ratio
    lm.diff_stats = diff_stats
    lm.stats = stats
    lm.trainer = _Trainer(precision)
    lm.register_buffer("border_mask", border_mask, persistent=False)
    lm.register_buffer("interior_mask", 1.0 - border_mask, persistent=False)
    lm.register_buffer("grid_static_features", grid_static_features, persistent=False)
    return lm
