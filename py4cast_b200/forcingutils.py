"""
External forcings (date / hour features and top-of-atmosphere solar radiation) generated on the GPU.

Mirrors reference py4cast/forcingutils.py — same function names, arguments and returned shapes:
`compute_day_of_years` (:19-33), `compute_hours_of_day` (:36-46),
`compute_seconds_from_start_of_year` (:49-58), `get_year_hour_forcing` (:61-90),
`generate_toa_radiation_forcing` (:93-132) — and `generate_forcings` of datasets/base.py:233-274.
The date arithmetic (a handful of scalars per sample) stays on the host exactly as in the
reference; the per-pixel part runs in ONE kernel (`p4c_forcing_fill`), which can also write the
five generated channels straight into their slots of a batch's forcing tensor
(`fill_external_forcings`) instead of building per-channel tensors that are expanded and
concatenated afterwards (base.py:509-515).  No CPU fallback: lat / lon must be CUDA tensors.
"""

from __future__ import annotations

import datetime as dt
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream
from .namedtensor import NamedTensor


def _valid_times(date: dt.datetime, output_terms: Sequence[dt.timedelta]) -> List[dt.datetime]:
    return [date + term for term in output_terms]


def compute_day_of_years(date: dt.datetime, output_terms: List[dt.timedelta]) -> np.ndarray:
    """Day of the year of each valid time (1st of January is 1)."""
    return np.asarray([t.timetuple().tm_yday for t in _valid_times(date, output_terms)])


def compute_hours_of_day(date: dt.datetime, output_terms: List[dt.timedelta]) -> np.ndarray:
    """Hour of the day of each valid time, minutes as a fraction."""
    return np.asarray([t.hour + t.minute / 60 for t in _valid_times(date, output_terms)])


def compute_seconds_from_start_of_year(date: dt.datetime, output_terms: List[dt.timedelta]) -> np.ndarray:
    """Seconds between the 1st of January of `date`'s year (not of the valid time's year: a window that
    crosses the new year keeps counting, as in the reference) and each valid time."""
    origin = dt.datetime(date.year, 1, 1)
    return np.asarray([(t - origin).total_seconds() for t in _valid_times(date, output_terms)])


def get_year_hour_forcing(date: dt.datetime, output_terms: List[dt.timedelta]) -> torch.Tensor:
    """(N_t, 4): sin / cos of the hour-of-day angle and of the time-of-year angle, mapped to [0, 1].
    A host tensor — four numbers per time slot — computed with the reference's float32 operation
    order (hours / 12 * pi, seconds / seconds_in_year * 2 * pi, (x + 1) / 2), so that it is bit-equal
    to it; the reference's year length rule (366 days iff year % 4 == 0) is kept."""
    seconds_in_year = (366 if date.year % 4 == 0 else 365) * 86400
    hour_angle = (torch.Tensor(compute_hours_of_day(date, output_terms)) / 12) * torch.pi
    year_angle = (torch.Tensor(compute_seconds_from_start_of_year(date, output_terms)) / seconds_in_year) * 2 * torch.pi
    columns = [fn(angle) for angle in (hour_angle, year_angle) for fn in (torch.sin, torch.cos)]
    return (torch.stack(columns, dim=1) + 1) / 2


def _solar_scalars(date_utc: dt.datetime, output_terms: Sequence[dt.timedelta]):
    """Per time slot: hour of day, sin / cos of the solar declination (float32, reference :101-117)."""
    day_of_years = compute_day_of_years(date_utc, list(output_terms))
    hours_of_day = compute_hours_of_day(date_utc, list(output_terms))
    dec = 23.45 * torch.sin(2 * np.pi * (284 + torch.Tensor(day_of_years)) / 365)
    dec_rad = dec * np.float32(np.pi / 180.0)
    return torch.Tensor(hours_of_day), torch.sin(dec_rad), torch.cos(dec_rad)


def _require_cuda(t: torch.Tensor, what: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.P4CError(f"py4cast_b200.forcingutils: {what} must be a CUDA tensor (no CPU fallback)")
    return t.to(torch.float32).contiguous()


def _fill(lat, lon, hour, sin_dec, cos_dec, dt4, out, ld: int, ch0: int) -> None:
    dev = lat.device
    hour, sin_dec, cos_dec = (v.to(dev, torch.float32).contiguous() for v in (hour, sin_dec, cos_dec))
    dt4 = None if dt4 is None else dt4.to(dev, torch.float32).contiguous()
    check(
        _lib.load().p4c_forcing_fill(
            ptr(lat), ptr(lon), ptr(hour), ptr(sin_dec), ptr(cos_dec), ptr(dt4), hour.numel(), lat.numel(),
            ptr(out), ld, ch0, stream(),
        ),
        "p4c_forcing_fill",
    )


def generate_toa_radiation_forcing(lat: torch.Tensor, lon: torch.Tensor, date_utc: dt.datetime,
                                   output_terms: List[dt.timedelta]) -> torch.Tensor:
    """Solar irradiation at the top of the atmosphere, (N_t, lat, lon, 1) on lat's device."""
    lat, lon = _require_cuda(lat, "lat"), _require_cuda(lon, "lon")
    hour, sd, cd = _solar_scalars(date_utc, output_terms)
    out = torch.empty(hour.numel(), *lat.shape, 1, dtype=torch.float32, device=lat.device)
    _fill(lat, lon, hour, sd, cd, None, out, 1, 0)
    return out


def generate_forcings(date: dt.datetime, timedeltas: List[dt.timedelta], grid) -> List[NamedTensor]:
    """The list of external-forcing NamedTensors of reference datasets/base.py:233-274
    (`grid` needs `.lat` / `.lon`, CUDA tensors of shape (lat, lon))."""
    datetime_forcing = get_year_hour_forcing(date, timedeltas).type(torch.float32).to(grid.lat.device)
    solar_forcing = generate_toa_radiation_forcing(grid.lat, grid.lon, date, timedeltas)
    return [
        NamedTensor(feature_names=["cos_hour", "sin_hour"], tensor=datetime_forcing[:, :2],
                    names=["timestep", "features"]),
        NamedTensor(feature_names=["cos_doy", "sin_doy"], tensor=datetime_forcing[:, 2:],
                    names=["timestep", "features"]),
        NamedTensor(feature_names=["toa_radiation"], tensor=solar_forcing,
                    names=["timestep", "lat", "lon", "features"]),
    ]


def external_forcing_scalars(dates: Sequence[dt.datetime], timedeltas: Sequence[dt.timedelta]) -> torch.Tensor:
    """The host part of the external forcings of a batch, for the input pipeline: a (B * T, 7) float32
    tensor [hour of day, sin / cos of the solar declination, cos_hour, sin_hour, cos_doy, sin_doy] per
    (sample, step) — 28 bytes per slot, computed by the data-loading workers with the reference's own date
    arithmetic (forcingutils.py:19-117) and shipped INSTEAD of five per-pixel channels."""
    rows = []
    for d in dates:
        h, sd, cd = _solar_scalars(d, timedeltas)
        dt4 = get_year_hour_forcing(d, list(timedeltas)).type(torch.float32)
        rows.append(torch.cat([h.unsqueeze(1), sd.unsqueeze(1), cd.unsqueeze(1), dt4], dim=1))
    return torch.cat(rows).to(torch.float32).contiguous()


def fill_external_forcings_from_scalars(forcing: torch.Tensor, ch0: int, lat: torch.Tensor, lon: torch.Tensor,
                                        scalars: torch.Tensor) -> None:
    """`fill_external_forcings` with the date arithmetic already done (`external_forcing_scalars`, on the
    device): one kernel launch, nothing on the host — what the training input path calls every step."""
    if not (forcing.is_cuda and forcing.dtype == torch.float32 and forcing.is_contiguous()):
        raise _lib.P4CError("fill_external_forcings: forcing must be a contiguous fp32 CUDA tensor")
    if not (scalars.is_cuda and scalars.dtype == torch.float32 and scalars.dim() == 2 and scalars.shape[1] == 7):
        raise _lib.P4CError("fill_external_forcings: scalars must be a (B * T, 7) fp32 CUDA tensor")
    F = forcing.shape[-1]
    slots = scalars.shape[0]
    if forcing.numel() != slots * lat.numel() * F:
        raise ValueError("fill_external_forcings: forcing / scalars / grid sizes do not match")
    cols = scalars.t().contiguous()  # (7, slots): hour, sin_dec, cos_dec, then the (slots, 4) block below
    dt4 = scalars[:, 3:7].contiguous()
    check(
        _lib.load().p4c_forcing_fill(ptr(lat), ptr(lon), ptr(cols[0]), ptr(cols[1]), ptr(cols[2]), ptr(dt4), slots,
                                     lat.numel(), ptr(forcing), F, ch0, stream()),
        "p4c_forcing_fill",
    )


def assemble_forcings(forcing: torch.Tensor, shipped: torch.Tensor, lat: torch.Tensor, lon: torch.Tensor,
                      scalars: torch.Tensor) -> None:
    """The whole forcing tensor of a batch in one pass: `forcing[..., :n] = shipped` (fp32 or bf16 -> fp32),
    `forcing[..., n:n+5]` = the five external channels (bit-identical with `fill_external_forcings_from_scalars`).
    Rows are written whole and in order, instead of a strided channel copy followed by 20-byte writes into
    84-byte rows.  Reference: datasets/base.py:503-515 (NWP forcings, then generate_forcings, concatenated)."""
    if not (forcing.is_cuda and forcing.dtype == torch.float32 and forcing.is_contiguous()):
        raise _lib.P4CError("assemble_forcings: forcing must be a contiguous fp32 CUDA tensor")
    if not (shipped.is_cuda and shipped.is_contiguous() and shipped.dtype in (torch.float32, torch.bfloat16)):
        raise _lib.P4CError("assemble_forcings: shipped channels must be a contiguous fp32 / bf16 CUDA tensor")
    if not (scalars.is_cuda and scalars.dtype == torch.float32 and scalars.dim() == 2 and scalars.shape[1] == 7):
        raise _lib.P4CError("assemble_forcings: scalars must be a (B * T, 7) fp32 CUDA tensor")
    F, n = forcing.shape[-1], shipped.shape[-1]
    slots = scalars.shape[0]
    if F != n + 5 or forcing.numel() != slots * lat.numel() * F or shipped.numel() != slots * lat.numel() * n:
        raise ValueError("assemble_forcings: forcing / shipped / scalars / grid sizes do not match")
    cols = scalars.t().contiguous()  # (7, slots): hour, sin_dec, cos_dec, ...
    dt4 = scalars[:, 3:7].contiguous()
    check(
        _lib.load().p4c_forcing_assemble(ptr(shipped), _lib.dtype_code(shipped.dtype), n, ptr(lat), ptr(lon),
                                         ptr(cols[0]), ptr(cols[1]), ptr(cols[2]), ptr(dt4), slots, lat.numel(),
                                         ptr(forcing), F, stream()),
        "p4c_forcing_assemble",
    )


def fill_external_forcings(forcing: torch.Tensor, ch0: int, lat: torch.Tensor, lon: torch.Tensor,
                           dates: Sequence[dt.datetime], timedeltas: Sequence[dt.timedelta]) -> None:
    """Writes [cos_hour, sin_hour, cos_doy, sin_doy, toa_radiation] of every (sample, step) into
    channels ch0..ch0+4 of a batch's forcing tensor (B, T, lat, lon, F_f) / (B, T, N, F_f) in ONE launch —
    what `generate_forcings` + `unsqueeze_and_expand_from_` + `NamedTensor.concat` produce per item
    in the reference (base.py:503-515), for a whole batch (`dates[b]` = reference date of sample b)."""
    lat, lon = _require_cuda(lat, "lat"), _require_cuda(lon, "lon")
    if not (forcing.is_cuda and forcing.dtype == torch.float32 and forcing.is_contiguous()):
        raise _lib.P4CError("fill_external_forcings: forcing must be a contiguous fp32 CUDA tensor")
    B, T, F = forcing.shape[0], forcing.shape[1], forcing.shape[-1]
    if len(dates) != B or len(timedeltas) != T or forcing.numel() != B * T * lat.numel() * F:
        raise ValueError("fill_external_forcings: shapes / dates / timedeltas do not match")
    hours, sds, cds, dt4s = [], [], [], []
    for d in dates:
        h, sd, cd = _solar_scalars(d, timedeltas)
        hours.append(h), sds.append(sd), cds.append(cd)
        dt4s.append(get_year_hour_forcing(d, list(timedeltas)).type(torch.float32))
    _fill(lat, lon, torch.cat(hours), torch.cat(sds), torch.cat(cds), torch.cat(dt4s), forcing, F, ch0)
