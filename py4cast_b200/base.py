"""
Plugin contract of the model registry: `ModelABC`, `ModelType`, layout helpers.

Mirror of the `mfai.pytorch.models.base` / `mfai.pytorch.models.utils` names the
reference imports (py4cast/lightning.py:16-21, py4cast/models.py:10-11,
py4cast_plugin_example.py:9,19-26).  mfai itself is an absent third-party
dependency (requirements.txt:26); the contract is taken from those call sites
and from doc/add_features_contribute.md:19-30.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from enum import Enum
from typing import Any, Tuple

import torch


class ModelType(Enum):
    GRAPH = 1
    CONVOLUTIONAL = 2
    VISION_TRANSFORMER = 3
    LLM = 4
    MULTIMODAL_LLM = 5


class ModelABC(ABC):
    """
    Contract every registered model honours (class attributes + `settings`).
    Concrete models also inherit torch.nn.Module.
    """

    # the plugin discovery registers only classes with register = True (models.py:36-41)
    register: bool = False

    @property
    @abstractmethod
    def onnx_supported(self) -> bool: ...

    @property
    @abstractmethod
    def settings_kls(self) -> Any: ...

    @property
    @abstractmethod
    def supported_num_spatial_dims(self) -> Tuple[int, ...]: ...

    @property
    @abstractmethod
    def num_spatial_dims(self) -> int: ...

    @property
    @abstractmethod
    def features_last(self) -> bool: ...

    @property
    @abstractmethod
    def model_type(self) -> ModelType: ...

    @property
    def features_second(self) -> bool:
        return not self.features_last

    def check_required_attributes(self) -> None:
        for attr in ("in_channels", "out_channels", "input_shape"):
            if not hasattr(self, attr):
                raise AttributeError(f"Missing required attribute : {attr}")


def expand_to_batch(x: torch.Tensor, batch_size: int) -> torch.Tensor:
    """(d1..dn) -> (B, d1..dn) without copying (lightning.py:296-300)."""
    return x.unsqueeze(0).expand(batch_size, *([-1] * x.dim()))


def features_last_to_second(x: torch.Tensor) -> torch.Tensor:
    """(B, *spatial, F) -> (B, F, *spatial)   (lightning.py:591-594)."""
    return x.movedim(-1, 1)


def features_second_to_last(x: torch.Tensor) -> torch.Tensor:
    """(B, F, *spatial) -> (B, *spatial, F)."""
    return x.movedim(1, -1)
