"""
ORACLE — TEST INFRASTRUCTURE ONLY.

Generates tests/golden/*.pt by running the reference's OWN code (py4cast/lightning.py and
py4cast/losses.py, loaded verbatim through oracle/refshim.py) on seeded inputs, on CPU, fp32.
Run in the build container (needs /root/reference):

    python -m oracle.make_golden

The reference has no golden vectors or known-answer tests for this path; these
fixtures are its outputs, produced here, and travel to the GPU box where /root/reference does
not exist.  Recipe: SURVEY.md §8(c) (seed 1234, N(0,1) fields, diff_std ~ U(0.2,1.5),
diff_mean ~ N(0,0.05), stats.std ~ U(0.5,5), border width 2-10).

Losses on GRAPH-shaped (B,T,N,F) tensors are evaluated by the reference on the equivalent
5-D (B,T,H,W,F) view, because losses.py:156,197 hard-code the feature dim as 4 and raise
IndexError on 4-D input.
"""

from __future__ import annotations

import os
import sys
from types import SimpleNamespace

import torch

from . import mfai_shim, refshim

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
NAMES5 = ["batch", "timestep", "lat", "lon", "features"]


class LinearTanh(torch.nn.Module):
    """Deterministic stand-in 'model': y = 0.5 * tanh(x W), features last or second."""

    def __init__(self, f_in, f_out, graph: bool, features_second: bool, seed: int):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.W = torch.nn.Parameter(torch.randn(f_in, f_out, generator=g) * 0.3)
        self.model_type = mfai_shim.ModelType.GRAPH if graph else mfai_shim.ModelType.CONVOLUTIONAL
        self.features_second = features_second

    def forward(self, x):
        if self.features_second:
            return (0.5 * torch.tanh(x.movedim(1, -1) @ self.W)).movedim(-1, 1)
        return 0.5 * torch.tanh(x @ self.W)


def make_problem(seed, B, T_in, T, H, W, F, Ff, Fs=4, border=2):
    g = torch.Generator().manual_seed(seed)
    fn = [f"w{i}" for i in range(F)]
    gn = [f"f{i}" for i in range(Ff)]
    p = SimpleNamespace()
    p.shape = dict(B=B, T_in=T_in, T=T, H=H, W=W, F=F, Ff=Ff, Fs=Fs, border=border)
    p.feature_names, p.forcing_names = fn, gn
    p.inputs = torch.randn(B, T_in, H, W, F, generator=g)
    p.forcing = torch.randn(B, T, H, W, Ff, generator=g)
    p.outputs = torch.randn(B, T, H, W, F, generator=g)
    x = torch.linspace(0, 1, W).view(1, W).expand(H, W)
    y = torch.linspace(0, 1, H).view(H, 1).expand(H, W)
    geo = torch.rand(H, W, generator=g)
    bm = torch.ones(H, W)
    bm[border:-border, border:-border] = 0.0
    p.statics = torch.stack([x, y, geo, bm], -1)[..., :Fs].contiguous()
    p.border_mask = bm.unsqueeze(-1)
    p.diff_std = torch.rand(F, generator=g) * 1.3 + 0.2
    p.diff_mean = torch.randn(F, generator=g) * 0.05
    p.stat_std = torch.rand(F, generator=g) * 4.5 + 0.5
    p.state_weight = torch.tensor([2.0 if i % 3 == 0 else 1.0 + i / 1000.0 for i in range(F)])
    return p


def dataset_info_for(p):
    stats = {n: {"std": p.stat_std[i], "mean": torch.tensor(0.0)} for i, n in enumerate(p.feature_names)}
    diff = {n: {"std": p.diff_std[i], "mean": p.diff_mean[i]} for i, n in enumerate(p.feature_names)}
    sw = {n: float(p.state_weight[i]) for i, n in enumerate(p.feature_names)}
    return SimpleNamespace(stats=refshim.DictStats(stats), diff_stats=refshim.DictStats(diff), state_weights=sw)


def run_reference_case(p, strategy, graph, features_second=False, num_inter_steps=1, model_seed=7):
    """Runs reference common_step + WeightedLoss(MSE) + ScaledLoss(MSE), returns tensors."""
    ref_losses, _ = refshim.load_reference()
    NT = mfai_shim.NamedTensor
    s = p.shape
    f_in = s["T_in"] * s["F"] + s["Fs"] + s["Ff"]
    model = LinearTanh(f_in, s["F"], graph, features_second, model_seed)
    info = dataset_info_for(p)
    border = p.border_mask.flatten(0, 1) if graph else p.border_mask
    statics = p.statics.flatten(0, 1) if graph else p.statics
    lm = refshim.make_reference_module(
        model, training_strategy=strategy, border_mask=border,
        grid_static_features=mfai_shim.expand_to_batch(statics, s["B"]),
        diff_stats=info.diff_stats, stats=info.stats, num_inter_steps=num_inter_steps,
    )
    batch = refshim.ItemBatch(
        inputs=NT(p.inputs.clone(), NAMES5, p.feature_names),
        forcing=NT(p.forcing.clone(), NAMES5, p.forcing_names),
        outputs=NT(p.outputs.clone(), NAMES5, p.feature_names),
    )
    pred, target = lm.common_step(batch, 0, "train")
    # the reference loss on the 5-D view
    pred5 = NT(pred.tensor.reshape(s["B"], s["T"], s["H"], s["W"], s["F"]), NAMES5, p.feature_names)
    tgt5 = NT(p.outputs.clone(), NAMES5, p.feature_names)
    holder = torch.nn.Module()
    wl = ref_losses.WeightedLoss("MSELoss", reduction="none")
    wl.prepare(holder, 1.0 - p.border_mask, info)
    mask = torch.ones_like(tgt5.tensor)  # lightning.py:797
    loss_bt = wl(pred5, tgt5, mask)
    batch_loss = torch.mean(loss_bt)  # lightning.py:816
    (dW,) = torch.autograd.grad(batch_loss, model.W, retain_graph=True)
    holder2 = torch.nn.Module()
    sl = ref_losses.ScaledLoss("MSELoss", reduction="none")
    sl.prepare(holder2, 1.0 - p.border_mask, info)
    scaled = sl(pred5, tgt5, mask)
    (dW_scaled,) = torch.autograd.grad(scaled.mean(), model.W)
    return {
        "pred": pred.tensor.detach().reshape(s["B"], s["T"], s["H"], s["W"], s["F"]).clone(),
        "weighted_mse_bt": loss_bt.detach().clone(),
        "batch_loss": batch_loss.detach().clone(),
        "dW": dW.clone(),
        "scaled_rmse_btf": scaled.detach().clone(),
        "dW_scaled": dW_scaled.clone(),
        "model_W": model.W.detach().clone(),
    }


def pack_problem(p):
    return {
        "shape": p.shape, "feature_names": p.feature_names, "forcing_names": p.forcing_names,
        "inputs": p.inputs, "forcing": p.forcing, "outputs": p.outputs, "statics": p.statics,
        "border_mask": p.border_mask, "diff_std": p.diff_std, "diff_mean": p.diff_mean,
        "stat_std": p.stat_std, "state_weight": p.state_weight,
    }


def loss_cases():
    """Reference losses on random 5-D tensors, with and without a partial mask, + d/dpred."""
    ref_losses, _ = refshim.load_reference()
    NT = mfai_shim.NamedTensor
    p = make_problem(4321, B=2, T_in=1, T=3, H=12, W=10, F=5, Ff=1, border=2)
    info = dataset_info_for(p)
    g = torch.Generator().manual_seed(99)
    pred0 = torch.randn(2, 3, 12, 10, 5, generator=g)
    tgt = torch.randn(2, 3, 12, 10, 5, generator=g)
    full = torch.ones_like(tgt)
    part = (torch.rand(2, 3, 12, 10, 5, generator=g) > 0.3).float()
    part[:, :, 5, 4:7, :] = 0.0  # three interior pixels masked everywhere -> union-mask denominator
    out = {"problem": pack_problem(p), "pred": pred0, "target": tgt, "mask_partial": part, "cases": {}}
    for kind in ("MSELoss", "L1Loss"):
        for mname, mask in (("full", full), ("partial", part)):
            for form in ("weighted", "weighted_map", "scaled"):
                pred = pred0.clone().requires_grad_(True)
                P, T = NT(pred, NAMES5, p.feature_names), NT(tgt, NAMES5, p.feature_names)
                holder = torch.nn.Module()
                if form == "scaled":
                    L = ref_losses.ScaledLoss(kind, reduction="none")
                    L.prepare(holder, 1.0 - p.border_mask, info)
                    val = L(P, T, mask)
                else:
                    L = ref_losses.WeightedLoss(kind, reduction="none")
                    L.prepare(holder, 1.0 - p.border_mask, info)
                    val = L(P, T, mask, reduce_spatial_dim=(form == "weighted"))
                gg = torch.Generator().manual_seed(5)
                gout = torch.rand(val.shape, generator=gg) + 0.5
                (dpred,) = torch.autograd.grad((val * gout).sum(), pred)
                out["cases"][f"{kind}/{mname}/{form}"] = {
                    "value": val.detach().clone(), "gout": gout, "dpred": dpred.clone(),
                }
    # CombinedLoss: 0.7 * WeightedLoss(MSE) + 0.3 * WeightedLoss(L1)
    pred = pred0.clone().requires_grad_(True)
    P, T = NT(pred, NAMES5, p.feature_names), NT(tgt, NAMES5, p.feature_names)
    conf = [
        {"class": "WeightedLoss", "weight": 0.7, "params": {"loss": "MSELoss", "reduction": "none"}},
        {"class": "WeightedLoss", "weight": 0.3, "params": {"loss": "L1Loss", "reduction": "none"}},
    ]
    cl = ref_losses.CombinedLoss(conf)
    cl.prepare(torch.nn.Module(), 1.0 - p.border_mask, info)
    val = cl(P, T, full)
    (dpred,) = torch.autograd.grad(val.mean(), pred)
    out["combined"] = {"config": conf, "value": val.detach().clone(), "dpred": dpred.clone()}
    return out


def main():
    if not refshim.reference_available():
        print("reference sources not available: cannot regenerate goldens", file=sys.stderr)
        return 1
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.manual_seed(1234)
    ar = {}
    small = dict(B=2, T_in=2, T=4, H=16, W=20, F=3, Ff=2, border=2)
    for name, kw in {
        "graph_scaled_ar": dict(strategy="scaled_ar", graph=True),
        "graph_diff_ar": dict(strategy="diff_ar", graph=True),
        "conv_scaled_ar": dict(strategy="scaled_ar", graph=False, features_second=True),
        "conv_diff_ar": dict(strategy="diff_ar", graph=False, features_second=True),
    }.items():
        p = make_problem(1234, **small)
        ar[name] = {"problem": pack_problem(p), "config": kw, "expected": run_reference_case(p, **kw)}
    p = make_problem(1235, B=2, T_in=1, T=3, H=16, W=20, F=3, Ff=2, border=2)
    kw = dict(strategy="scaled_ar", graph=True, num_inter_steps=2)
    ar["graph_scaled_ar_inter2"] = {"problem": pack_problem(p), "config": kw,
                                    "expected": run_reference_case(p, **kw)}
    p = make_problem(1236, B=1, T_in=3, T=5, H=12, W=12, F=4, Ff=3, border=3)
    kw = dict(strategy="scaled_ar", graph=True)
    ar["graph_scaled_ar_tin3"] = {"problem": pack_problem(p), "config": kw,
                                  "expected": run_reference_case(p, **kw)}
    # BASELINE config #1 shape: 64x64, F=4, B=2, T_in=1, T=1, scaled_ar (conv layout)
    p = make_problem(1237, B=2, T_in=1, T=1, H=64, W=64, F=4, Ff=2, border=10)
    kw = dict(strategy="scaled_ar", graph=False, features_second=True)
    ar["config1_conv_64x64"] = {"problem": pack_problem(p), "config": kw,
                                "expected": run_reference_case(p, **kw)}
    torch.save(ar, os.path.join(GOLDEN_DIR, "ar_rollout.pt"))
    torch.save(loss_cases(), os.path.join(GOLDEN_DIR, "losses.pt"))
    for f in sorted(os.listdir(GOLDEN_DIR)):
        print(f, os.path.getsize(os.path.join(GOLDEN_DIR, f)), "bytes")
    return 0


if __name__ == "__main__":
    sys.exit(main())
