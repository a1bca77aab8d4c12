"""
ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under py4cast_b200/ may import this.

Minimal stand-ins for the parts of the third-party `mfai` package (pinned by
the reference at mfai v5.0.1, requirements.txt:26, absent from /root/reference
and from this image) that the reference's hot-path files import:

  * mfai.pytorch.namedtensor.NamedTensor   (lightning.py:27 via datasets.base, losses.py:11)
  * mfai.pytorch.models.base.{ModelABC, ModelType}      (lightning.py:16, models.py:11)
  * mfai.pytorch.models.utils.{expand_to_batch, features_last_to_second,
    features_second_to_last}                            (lightning.py:17-21)

Written independently from py4cast_b200/namedtensor.py on purpose: the
verbatim reference code is executed on top of THIS shim to produce golden
vectors, so a bug in the product's NamedTensor cannot leak into the oracle.
Only the methods the hot path calls are provided.
"""

from __future__ import annotations

from enum import Enum

import torch


class NamedTensor:
    def __init__(self, tensor, names, feature_names, feature_dim_name="features"):
        assert tensor.dim() == len(names), (tensor.shape, names)
        assert tensor.shape[list(names).index(feature_dim_name)] == len(feature_names)
        self.tensor = tensor
        self.names = list(names)
        self.feature_names = list(feature_names)
        self.feature_dim_name = feature_dim_name

    _SPATIAL = ("lat", "lon", "ngrid")

    @property
    def device(self):
        return self.tensor.device

    @property
    def spatial_dim_idx(self):
        return sorted(i for i, n in enumerate(self.names) if n in self._SPATIAL)

    @property
    def num_spatial_dims(self):
        return len(self.spatial_dim_idx)

    def dim_index(self, name):
        return self.names.index(name)

    def dim_size(self, name):
        return self.tensor.shape[self.names.index(name)] if name in self.names else None

    def flatten_(self, new_name, start, end):
        self.tensor = self.tensor.flatten(start, end)
        self.names = self.names[:start] + [new_name] + self.names[end + 1 :]

    def select_tensor_dim(self, name, idx):
        return self.tensor.select(self.names.index(name), idx)

    def select_dim(self, name, idx):
        return NamedTensor(
            self.select_tensor_dim(name, idx),
            [n for n in self.names if n != name],
            self.feature_names,
            self.feature_dim_name,
        )

    def index_select_tensor_dim(self, name, indices):
        idx = torch.as_tensor(list(indices), dtype=torch.long, device=self.tensor.device)
        return self.tensor.index_select(self.names.index(name), idx)

    def clone(self):
        return NamedTensor(
            self.tensor.clone().detach(), list(self.names), list(self.feature_names)
        )

    def __getitem__(self, feature):
        i = self.feature_names.index(feature)
        return self.tensor.narrow(self.names.index(self.feature_dim_name), i, 1)

    @classmethod
    def new_like(cls, tensor, other):
        return cls(tensor, other.names, other.feature_names, other.feature_dim_name)

    @classmethod
    def expand_to_batch_like(cls, tensor, other):
        return cls(tensor, ["batch"] + other.names, other.feature_names)


class ModelType(Enum):
    GRAPH = 1
    CONVOLUTIONAL = 2
    VISION_TRANSFORMER = 3
    LLM = 4
    MULTIMODAL_LLM = 5


class ModelABC:
    register = False

    @property
    def features_second(self):
        return not self.features_last

    def check_required_attributes(self):
        pass


def expand_to_batch(x: torch.Tensor, batch_size: int) -> torch.Tensor:
    """(d1..dn) -> (B, d1..dn) view."""
    return x.unsqueeze(0).expand(batch_size, *([-1] * x.dim()))


def features_last_to_second(x: torch.Tensor) -> torch.Tensor:
    return x.movedim(-1, 1)


def features_second_to_last(x: torch.Tensor) -> torch.Tensor:
    return x.movedim(1, -1)
