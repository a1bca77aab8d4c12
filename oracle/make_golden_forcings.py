"""Golden vectors for the external forcings: runs the VERBATIM reference
/root/reference/py4cast/forcingutils.py (pure numpy / torch, imported unmodified) on a small
Titan-like lat / lon grid and writes tests/golden/forcings.pt.  Run in the build container only."""
import datetime as dt
import importlib.util
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    spec = importlib.util.spec_from_file_location("ref_forcingutils", "/root/reference/py4cast/forcingutils.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    g = torch.Generator().manual_seed(1234)
    H, W = 24, 40
    lat = torch.linspace(55.4, 37.5, H).view(H, 1).expand(H, W).contiguous() + 0.01 * torch.randn(H, W, generator=g)
    lon = torch.linspace(-12.0, 16.0, W).view(1, W).expand(H, W).contiguous() + 0.01 * torch.randn(H, W, generator=g)
    cases = {}
    for name, date, step_h, n in (
        ("summer_noon", dt.datetime(2023, 6, 21, 12, 0), 1, 6),
        ("leap_year_night", dt.datetime(2024, 2, 28, 21, 30), 3, 9),
        ("new_year", dt.datetime(2022, 12, 31, 18, 0), 6, 5),
    ):
        terms = [dt.timedelta(hours=step_h * (i + 1)) for i in range(n)]
        f4 = ref.get_year_hour_forcing(date, terms).type(torch.float32)
        toa = ref.generate_toa_radiation_forcing(lat, lon, date, terms).type(torch.float32)
        cases[name] = {"date": date, "terms": terms, "datetime_forcing": f4, "toa": toa}
    torch.save({"lat": lat, "lon": lon, "cases": cases}, os.path.join(ROOT, "tests", "golden", "forcings.pt"))
    for k, v in cases.items():
        print(k, tuple(v["datetime_forcing"].shape), tuple(v["toa"].shape), float(v["toa"].max()))


if __name__ == "__main__":
    main()
