"""Golden vectors for the validation metrics: loads the VERBATIM reference
/root/reference/py4cast/metrics.py under import stubs (torchmetrics.Metric, py4cast.datasets.base,
py4cast.plots are absent / irrelevant here) and writes tests/golden/metrics.pt.  Build container only."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_reference_metrics():
    class Metric:  # the two members the reference's metric classes use
        def __init__(self):
            self.device = torch.device("cpu")

        def add_state(self, name, default, dist_reduce_fx=None):
            setattr(self, name, default)

        def reset(self):
            pass

    stubs = {
        "torchmetrics": types.SimpleNamespace(Metric=Metric),
        "py4cast": types.ModuleType("py4cast"),
        "py4cast.datasets": types.ModuleType("py4cast.datasets"),
        "py4cast.datasets.base": types.SimpleNamespace(DatasetInfo=object, NamedTensor=object),
        "py4cast.plots": types.SimpleNamespace(plot_log_psd=lambda *a, **k: None),
    }
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location("ref_metrics", "/root/reference/py4cast/metrics.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


class _NT:  # what MetricACC.update touches of a NamedTensor
    def __init__(self, tensor, feature_names, spatial_dim_idx):
        self.tensor, self.feature_names, self.spatial_dim_idx = tensor, feature_names, spatial_dim_idx


def main():
    ref = load_reference_metrics()
    g = torch.Generator().manual_seed(4321)
    out = {"psd": {}, "acc": {}}
    for name, (B, C, H, W) in {"square": (3, 2, 32, 32), "wide": (2, 3, 24, 40), "tall": (2, 1, 48, 20)}.items():
        x = torch.randn(B, C, H, W, generator=g, dtype=torch.float64).numpy()
        out["psd"][name] = {"x": torch.from_numpy(x), "psd": torch.from_numpy(ref.power_spectral_density(x))}
    # MetricACC.update on a graph-shaped and a grid-shaped batch
    for name, shape, sdims in (("graph", (2, 3, 96, 4), [2]), ("grid", (2, 2, 8, 12, 3), [2, 3])):
        pred = torch.randn(*shape, generator=g)
        tgt = pred + 0.3 * torch.randn(*shape, generator=g)
        mask = (torch.rand(*shape, generator=g) > 0.1).float()
        means = torch.randn(shape[-1], generator=g) * 0.2
        m = ref.MetricACC.__new__(ref.MetricACC)
        m.device = torch.device("cpu")
        m.climate_means = means
        m.sum_acc = torch.tensor(0.0)
        m.step_count = torch.tensor(0.0)
        fn = [f"w{i}" for i in range(shape[-1])]
        m.update(_NT(pred, fn, sdims), _NT(tgt, fn, sdims), mask)
        out["acc"][name] = {"pred": pred, "target": tgt, "mask": mask, "climate_means": means,
                            "spatial_dims": sdims, "acc": m.sum_acc.clone()}
    torch.save(out, os.path.join(ROOT, "tests", "golden", "metrics.pt"))
    for k, v in out["psd"].items():
        print("psd", k, tuple(v["psd"].shape), bool(np.isfinite(v["psd"].numpy()).all()))
    for k, v in out["acc"].items():
        print("acc", k, tuple(v["acc"].shape))


if __name__ == "__main__":
    main()
