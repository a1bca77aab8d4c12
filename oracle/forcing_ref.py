"""
TEST INFRASTRUCTURE ONLY (oracle): CPU restatement of the reference's external-forcing generation,
py4cast/forcingutils.py:61-132 + datasets/base.py:233-274, 503-515.  Pinned against the verbatim
reference by tests/golden/forcings.pt (oracle/make_golden_forcings.py imports
/root/reference/py4cast/forcingutils.py unmodified).  Only tests/ may import this module.
"""
import datetime as dt
from typing import List

import numpy as np
import torch


def year_hour_forcing(date: dt.datetime, terms: List[dt.timedelta]) -> torch.Tensor:
    """forcingutils.py:61-90 -> (T, 4) in [0, 1]."""
    hours = np.asarray([(date + t).hour + (date + t).minute / 60 for t in terms])  # :36-46
    start = dt.datetime(date.year, 1, 1)
    secs = np.asarray([(date + t - start).total_seconds() for t in terms])  # :49-58
    seconds_in_year = (366 if date.year % 4 == 0 else 365) * 24 * 60 * 60  # :71-72
    hour_angle = (torch.Tensor(hours) / 12) * torch.pi  # :74
    year_angle = (torch.Tensor(secs) / seconds_in_year) * 2 * torch.pi  # :75-77
    f = torch.stack((torch.sin(hour_angle), torch.cos(hour_angle), torch.sin(year_angle), torch.cos(year_angle)), dim=1)
    return (f + 1) / 2  # :88


def toa_radiation(lat: torch.Tensor, lon: torch.Tensor, date: dt.datetime, terms: List[dt.timedelta]) -> torch.Tensor:
    """forcingutils.py:93-132 -> (T, H, W, 1)."""
    days, hours = [], []
    for t in terms:
        d = date + t
        days.append((d - dt.datetime(d.year, 1, 1)).days + 1)  # :19-33
        hours.append(d.hour + d.minute / 60)
    rad = np.float32(np.pi / 180.0)
    hours_lcl = torch.Tensor(np.asarray(hours)).unsqueeze(-1).unsqueeze(-1) + lon / 15  # :106
    omega_rad = (15 * (hours_lcl - 12)) * rad  # :107-108
    dec = 23.45 * torch.sin(2 * np.pi * (284 + torch.Tensor(np.asarray(days))) / 365).unsqueeze(-1).unsqueeze(-1)  # :116-118
    dec_rad = dec * rad
    phi_rad = torch.Tensor(lat) * rad  # :123-124
    cos_sza = torch.sin(phi_rad) * torch.sin(dec_rad) + torch.cos(phi_rad) * torch.cos(dec_rad) * torch.cos(omega_rad)  # :127-129
    return torch.fmax(torch.tensor(0), 1366 * cos_sza).unsqueeze(-1)  # :132


def external_forcing_block(lat, lon, date, terms) -> torch.Tensor:
    """(T, H, W, 5): [cos_hour, sin_hour, cos_doy, sin_doy, toa_radiation] broadcast over the grid and
    concatenated in the order of base.py:248-272 / :509-515."""
    T = len(terms)
    f4 = year_hour_forcing(date, terms).type(torch.float32)
    toa = toa_radiation(lat, lon, date, terms).type(torch.float32)
    return torch.cat([f4.view(T, 1, 1, 4).expand(T, *lat.shape, 4), toa], dim=-1)
