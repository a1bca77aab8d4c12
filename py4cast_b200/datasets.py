"""
Batch containers either side of the AR step: `Item`, `ItemBatch`, `collate_fn`, `Statics`,
`Stats`, `DatasetInfo` — the input contract of `common_step`.

Mirror of reference py4cast/datasets/base.py:40-230 (Item / ItemBatch / collate_fn / Statics),
:277-300 (DatasetInfo fields) and py4cast/datasets/access.py:355-390 (Stats.to_list), restricted
to what the hot path reads.  Disk datasets, GRIB access and forcing generation are out of scope
(SURVEY.md §2 rows 10-15); `synthetic_dataset_info` / `synthetic_batch` provide Titan-shaped
random data (titan.yaml:30-77: 512x640 grid, 21 weather features, 21 forcings, 4 statics).
"""

from __future__ import annotations

import datetime as dt
from dataclasses import dataclass, field, fields
from functools import cached_property
from typing import Dict, List, Literal, Optional, Tuple, Union

import torch

from .namedtensor import NamedTensor


@dataclass(slots=True)
class Item:
    """inputs / outputs / forcing: (timestep, lat, lon, features)."""

    inputs: Optional[NamedTensor]
    forcing: Optional[NamedTensor]
    outputs: NamedTensor
    validity_times: Optional[List[dt.datetime]] = None

    def unsqueeze_(self, dim_name: str, dim_index: int):
        for nt in (self.outputs, self.inputs, self.forcing):
            if nt:
                nt.unsqueeze_(dim_name, dim_index)

    def squeeze_(self, dim_name: Union[List[str], str]):
        for nt in (self.outputs, self.inputs, self.forcing):
            if nt:
                nt.squeeze_(dim_name)

    def to_(self, *args, **kwargs):
        for nt in (self.outputs, self.inputs, self.forcing):
            if nt:
                nt.to_(*args, **kwargs)

    def pin_memory(self):
        for nt in (self.outputs, self.inputs, self.forcing):
            if nt:
                nt.pin_memory_()
        return self

    def __post_init__(self):
        if self.inputs is not None and self.outputs is not None:
            if self.inputs.names != self.outputs.names:
                raise ValueError(
                    f"Inputs and outputs must have the same dim names, got {self.inputs.names} and {self.outputs.names}"
                )
            if self.inputs.feature_names != self.outputs.feature_names:
                raise ValueError(
                    f"Inputs and outputs must have the same feature names, "
                    f"got {self.inputs.feature_names} and {self.outputs.feature_names}"
                )


@dataclass
class ItemBatch(Item):
    """inputs / outputs / forcing: (batch, timestep, lat, lon, features)."""

    @cached_property
    def batch_size(self):
        ref = self.outputs if self.outputs is not None else self.forcing
        return ref.dim_size("batch")

    @cached_property
    def num_input_steps(self):
        if self.inputs.dim_size("timestep") is None:
            return self.outputs.dim_size("timestep")
        return self.inputs.dim_size("timestep")

    @cached_property
    def num_pred_steps(self):
        ref = self.outputs if self.outputs is not None else self.forcing
        return ref.dim_size("timestep")


def collate_fn(items: List[Item]) -> ItemBatch:
    """Stack items into a batch, cast to float32, prepend the "batch" dim (base.py:173-195)."""
    batch_of_items = {}
    for field_name in (f.name for f in fields(Item)):
        if field_name == "validity_times":
            batch_of_items[field_name] = [getattr(item, field_name) for item in items]
        else:
            stacked = torch.stack([getattr(item, field_name).tensor for item in items]).type(torch.float32)
            batch_of_items[field_name] = NamedTensor.expand_to_batch_like(stacked, getattr(items[0], field_name))
    return ItemBatch(**batch_of_items)


@dataclass
class Statics:
    """Static fields; `grid_statics` holds x, y, geopotential, border_mask (access.py:303)."""

    grid_statics: NamedTensor
    grid_shape: Tuple[int, int]
    border_mask: torch.Tensor = field(init=False)
    interior_mask: torch.Tensor = field(init=False)

    def __post_init__(self):
        self.border_mask = self.grid_statics["border_mask"]
        self.interior_mask = 1.0 - self.border_mask

    @cached_property
    def meshgrid(self) -> torch.Tensor:
        """(2, H, W) tensor of x, y."""
        return torch.cat([self.grid_statics["x"], self.grid_statics["y"]], dim=-1).permute(2, 0, 1)

    def register_buffers(self, lm: torch.nn.Module, persistent: bool = False):
        """Registers the tensor fields as buffers (utils.py:74-89 RegisterFieldsMixin)."""
        for f in fields(self):
            v = getattr(self, f.name)
            if isinstance(v, torch.Tensor):
                lm.register_buffer(f.name, v, persistent=persistent)


class Stats:
    """name -> {"mean","std","min","max"} 0-d tensors; file-backed or in-memory (access.py:355-390)."""

    def __init__(self, fname=None, stats: Optional[dict] = None):
        if stats is None:
            stats = torch.load(fname, "cpu", weights_only=True)
        self.stats = stats
        self.fname = fname

    def items(self):
        return self.stats.items()

    def __getitem__(self, shortname: str):
        return self.stats[shortname]

    def to_list(self, stat_name: Literal["mean", "std", "min", "max"], shortnames: List[str],
                dtype: torch.dtype = torch.float32):
        if len(shortnames) > 0:
            return torch.stack(
                [torch.as_tensor(self[name][stat_name]) for name in shortnames], dim=0
            ).type(dtype)
        return []


@dataclass(slots=True)
class DatasetInfo:
    """Fields of reference base.py:277-300 that the training path reads."""

    name: str
    domain_info: object
    units: Dict[str, str]
    weather_dim: int
    forcing_dim: int
    pred_step: dt.timedelta
    statics: Statics
    stats: Stats
    diff_stats: Stats
    state_weights: Dict[str, float]
    shortnames: Dict[str, List[str]] = None


# --------------------------------------------------------------------- synthetic data
def make_statics(H: int, W: int, border_size: int = 10, seed: int = 0) -> Statics:
    """x, y in [0,1], geopotential in [0,1], border mask of `border_size` pixels
    (access.py:227-237, 267-308)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.linspace(0.0, 1.0, W).view(1, W).expand(H, W)
    y = torch.linspace(0.0, 1.0, H).view(H, 1).expand(H, W)
    geo = torch.rand(H, W, generator=g)
    geo = (geo - geo.min()) / (geo.max() - geo.min())
    border = torch.ones(H, W)
    if border_size > 0:
        border[border_size:-border_size, border_size:-border_size] = 0.0
    else:
        border.zero_()
    t = torch.stack([x, y, geo, border], dim=-1).to(torch.float32).contiguous()
    nt = NamedTensor(t, ["lat", "lon", "features"], ["x", "y", "geopotential", "border_mask"])
    return Statics(grid_statics=nt, grid_shape=(H, W))


def synthetic_dataset_info(H: int = 512, W: int = 640, num_features: int = 21, num_forcing: int = 21,
                           border_size: int = 10, seed: int = 1234) -> DatasetInfo:
    """Titan-shaped synthetic dataset description (SURVEY.md §8c golden-vector recipe)."""
    g = torch.Generator().manual_seed(seed)
    names = [f"w{i:02d}" for i in range(num_features)]
    stats, diff_stats, weights = {}, {}, {}
    for i, n in enumerate(names):
        stats[n] = {
            "mean": torch.randn((), generator=g) * 0.1,
            "std": torch.rand((), generator=g) * 4.5 + 0.5,
            "min": torch.tensor(-10.0), "max": torch.tensor(10.0),
        }
        diff_stats[n] = {
            "mean": torch.randn((), generator=g) * 0.05,
            "std": torch.rand((), generator=g) * 1.3 + 0.2,
        }
        weights[n] = 2.0 if i < 5 else 1.0 + (i % 4) / 1000.0  # titan/__init__.py:23-30 flavour
    return DatasetInfo(
        name="synthetic_titan", domain_info=None, units={n: "1" for n in names},
        weather_dim=num_features, forcing_dim=num_forcing, pred_step=dt.timedelta(hours=1),
        statics=make_statics(H, W, border_size), stats=Stats(stats=stats),
        diff_stats=Stats(stats=diff_stats), state_weights=weights,
        shortnames={"input": [], "input_output": names, "output": [],
                    "forcing": [f"f{i:02d}" for i in range(num_forcing)]},
    )


def synthetic_batch(info: DatasetInfo, batch_size: int, num_input_steps: int, num_pred_steps: int,
                    seed: int = 0, device="cpu", pin: bool = False) -> ItemBatch:
    """Standardised N(0,1) fields with the ItemBatch layout (B, T, lat, lon, F) float32."""
    H, W = info.statics.grid_shape
    g = torch.Generator().manual_seed(seed)
    names = ["batch", "timestep", "lat", "lon", "features"]
    wn = info.shortnames["input_output"]
    fn = info.shortnames["forcing"]

    def mk(T, F):
        t = torch.randn(batch_size, T, H, W, F, generator=g, dtype=torch.float32)
        if pin:
            t = t.pin_memory()
        return t.to(device)

    return ItemBatch(
        inputs=NamedTensor(mk(num_input_steps, len(wn)), names, wn),
        forcing=NamedTensor(mk(num_pred_steps, len(fn)), names, fn),
        outputs=NamedTensor(mk(num_pred_steps, len(wn)), names, wn),
    )
