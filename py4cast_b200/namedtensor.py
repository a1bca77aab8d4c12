"""
NamedTensor: a tensor + dimension names + feature names.

Host-side mirror of the `mfai.pytorch.namedtensor.NamedTensor` API that the
reference hot path touches (mfai is a third-party dependency absent from
/root/reference; the API surface is pinned by its call sites:
reference py4cast/lightning.py:526-676, py4cast/losses.py:130-210,
py4cast/datasets/base.py:40-230, tests/test_datasets.py:15-111 and
doc/features.md:134-164).  Written from the call sites, not from mfai source.

Pure metadata bookkeeping: no kernels live here.
"""

from __future__ import annotations

from copy import deepcopy
from typing import Any, Iterable, List, Sequence, Union

import torch

SPATIAL_DIM_NAMES = ("lat", "lon", "ngrid")


class NamedTensor:
    """Wrapper linking a torch.Tensor with dim names and per-feature names."""

    SPATIAL_DIM_NAMES = SPATIAL_DIM_NAMES

    def __init__(
        self,
        tensor: torch.Tensor,
        names: Sequence[str],
        feature_names: Sequence[str],
        feature_dim_name: str = "features",
    ):
        names = list(names)
        feature_names = list(feature_names)
        if len(tensor.shape) != len(names):
            raise ValueError(
                f"Number of names ({len(names)}) must match number of dimensions "
                f"({len(tensor.shape)})"
            )
        if feature_dim_name not in names:
            raise ValueError(f"Dimension '{feature_dim_name}' must be in names {names}")
        if tensor.shape[names.index(feature_dim_name)] != len(feature_names):
            raise ValueError(
                f"Number of feature names ({len(feature_names)}) must match the "
                f"'{feature_dim_name}' dimension size "
                f"({tensor.shape[names.index(feature_dim_name)]})"
            )
        self.tensor = tensor
        self.names = names
        self.feature_names = feature_names
        self.feature_dim_name = feature_dim_name

    # ------------------------------------------------------------------ meta
    @property
    def device(self) -> torch.device:
        return self.tensor.device

    @property
    def ndims(self) -> int:
        return len(self.names)

    @property
    def num_spatial_dims(self) -> int:
        return len([n for n in self.names if n in SPATIAL_DIM_NAMES])

    @property
    def spatial_dim_idx(self) -> List[int]:
        return sorted(self.names.index(n) for n in self.names if n in SPATIAL_DIM_NAMES)

    @property
    def feature_names_to_idx(self) -> dict:
        return {name: i for i, name in enumerate(self.feature_names)}

    def dim_index(self, dim_name: str) -> int:
        return self.names.index(dim_name)

    def dim_size(self, dim_name: str) -> int | None:
        try:
            return self.tensor.size(self.names.index(dim_name))
        except ValueError:
            return None

    # ------------------------------------------------------------ selection
    def __getitem__(self, feature_name: str) -> torch.Tensor:
        """Tensor of one feature, feature dim kept with size 1."""
        try:
            idx = self.feature_names_to_idx[feature_name]
        except KeyError as e:
            raise ValueError(
                f"Feature {feature_name} not found in {self.feature_names}"
            ) from e
        return self.tensor.narrow(self.dim_index(self.feature_dim_name), idx, 1)

    def select_tensor_dim(self, dim_name: str, index: int) -> torch.Tensor:
        return self.tensor.select(self.dim_index(dim_name), index)

    def index_select_tensor_dim(
        self, dim_name: str, indices: Union[Iterable[int], torch.Tensor]
    ) -> torch.Tensor:
        if not isinstance(indices, torch.Tensor):
            indices = torch.tensor(list(indices), dtype=torch.long)
        return self.tensor.index_select(
            self.dim_index(dim_name), indices.to(self.tensor.device)
        )

    def select_dim(self, dim_name: str, index: int) -> "NamedTensor":
        if dim_name == self.feature_dim_name:
            raise ValueError("select_dim on the feature dimension is not allowed")
        return NamedTensor(
            self.select_tensor_dim(dim_name, index),
            [n for n in self.names if n != dim_name],
            list(self.feature_names),
            self.feature_dim_name,
        )

    def index_select_dim(self, dim_name: str, indices) -> "NamedTensor":
        t = self.index_select_tensor_dim(dim_name, indices)
        feats = (
            [self.feature_names[i] for i in indices]
            if dim_name == self.feature_dim_name
            else list(self.feature_names)
        )
        return NamedTensor(t, list(self.names), feats, self.feature_dim_name)

    def iter_dim(self, dim_name: str):
        for i in range(self.dim_size(dim_name)):
            yield self.select_dim(dim_name, i)

    def iter_tensor_dim(self, dim_name: str):
        for i in range(self.dim_size(dim_name)):
            yield self.select_tensor_dim(dim_name, i)

    # ------------------------------------------------------------- reshapes
    def flatten_(self, flatten_dim_name: str, start_dim: int, end_dim: int) -> None:
        """Flatten [start_dim, end_dim] in place into one dim called flatten_dim_name."""
        self.tensor = torch.flatten(self.tensor, start_dim, end_dim)
        self.names = (
            self.names[:start_dim] + [flatten_dim_name] + self.names[end_dim + 1 :]
        )

    def unflatten_(self, dim: int, unflattened_size, unflatten_dim_name) -> None:
        self.tensor = self.tensor.unflatten(dim, tuple(unflattened_size))
        self.names = self.names[:dim] + list(unflatten_dim_name) + self.names[dim + 1 :]

    def unsqueeze_(self, dim_name: str, dim_index: int) -> None:
        self.tensor = torch.unsqueeze(self.tensor, dim_index)
        self.names.insert(dim_index, dim_name)

    def squeeze_(self, dim_name: Union[List[str], str]) -> None:
        if isinstance(dim_name, str):
            dim_name = [dim_name]
        for name in dim_name:
            self.tensor = torch.squeeze(self.tensor, dim=self.names.index(name))
            self.names.remove(name)

    def rearrange_(self, einops_str: str) -> None:
        import einops

        self.tensor = einops.rearrange(self.tensor, einops_str)
        self.names = einops_str.split("->")[1].strip().split(" ")

    def unsqueeze_and_expand_from_(self, other: "NamedTensor") -> None:
        """Add (and expand along) every dim of `other` that self lacks."""
        for i, name in enumerate(other.names):
            if name not in self.names:
                self.unsqueeze_(name, i)
                shape = list(self.tensor.shape)
                shape[i] = other.tensor.shape[i]
                self.tensor = self.tensor.expand(*shape)

    # ------------------------------------------------------------ in-place
    def to_(self, *args, **kwargs) -> None:
        self.tensor = self.tensor.to(*args, **kwargs)

    def type_(self, new_type) -> None:
        self.tensor = self.tensor.type(new_type)

    def pin_memory_(self) -> None:
        self.tensor = self.tensor.pin_memory()

    def clone(self) -> "NamedTensor":
        return NamedTensor(
            self.tensor.clone().detach(),
            deepcopy(self.names),
            deepcopy(self.feature_names),
            self.feature_dim_name,
        )

    # --------------------------------------------------------- constructors
    @classmethod
    def new_like(cls, tensor: torch.Tensor, other: "NamedTensor") -> "NamedTensor":
        """New NamedTensor with `other`'s names around new data."""
        return cls(
            tensor, list(other.names), list(other.feature_names), other.feature_dim_name
        )

    @classmethod
    def expand_to_batch_like(
        cls, tensor: torch.Tensor, other: "NamedTensor"
    ) -> "NamedTensor":
        """`tensor` has one more leading dim ("batch") than `other`."""
        names = ["batch"] + list(other.names)
        if tensor.dim() != len(names):
            raise ValueError(
                f"Tensor dim {tensor.dim()} must be len(other.names)+1 = {len(names)}"
            )
        return cls(tensor, names, list(other.feature_names), other.feature_dim_name)

    @staticmethod
    def concat(nts: List["NamedTensor"]) -> "NamedTensor":
        """Concatenate along the feature dim; all other dims/names must agree."""
        if len(nts) == 0:
            raise ValueError("Cannot concatenate an empty list of NamedTensors")
        if len(nts) == 1:
            return nts[0].clone()
        first = nts[0]
        for nt in nts[1:]:
            if nt.names != first.names:
                raise ValueError(
                    f"NamedTensors must have the same dimension names to concatenate, "
                    f"got {first.names} and {nt.names}"
                )
        all_feats = [f for nt in nts for f in nt.feature_names]
        if len(set(all_feats)) != len(all_feats):
            raise ValueError(f"Duplicate feature names in concatenation: {all_feats}")
        fdim = first.dim_index(first.feature_dim_name)
        return NamedTensor(
            torch.cat([nt.tensor for nt in nts], dim=fdim),
            list(first.names),
            all_feats,
            first.feature_dim_name,
        )

    @staticmethod
    def stack(nts: List["NamedTensor"], dim_name: str, dim: int) -> "NamedTensor":
        first = nts[0]
        names = list(first.names)
        names.insert(dim, dim_name)
        return NamedTensor(
            torch.stack([nt.tensor for nt in nts], dim=dim),
            names,
            list(first.feature_names),
            first.feature_dim_name,
        )

    def __or__(self, other: Any) -> "NamedTensor":
        if other is None:
            return self
        if not isinstance(other, NamedTensor):
            raise ValueError(f"Cannot concatenate NamedTensor with {type(other)}")
        return NamedTensor.concat([self, other])

    def __ror__(self, other: Any) -> "NamedTensor":
        return self.__or__(other)

    # --------------------------------------------------------------- repr
    def __str__(self) -> str:
        head = (
            f"--- NamedTensor ---\nNames: {self.names}\n"
            f"Tensor Shape: {self.tensor.shape})\nFeatures:\n"
        )
        rows = []
        for name in self.feature_names:
            t = self[name]
            rows.append([name, float(t.min()), float(t.max())])
        try:
            from tabulate import tabulate

            body = tabulate(
                rows, headers=["Feature name", "Min", "Max"], tablefmt="simple_outline"
            )
        except Exception:  # pragma: no cover
            body = "\n".join(f"{r[0]}: [{r[1]}, {r[2]}]" for r in rows)
        return head + body

    __repr__ = __str__
