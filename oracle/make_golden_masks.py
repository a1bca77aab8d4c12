"""
ORACLE — TEST INFRASTRUCTURE ONLY.

Golden vectors for the NaN-mask (`mask_on_nan`), masked-auto-encoder (`mask_ratio`) and
`downscaling_only` branches of the AR step, produced by the reference's OWN code (py4cast/lightning.py:568-569, 588-589, 600-601,
732-797 and py4cast/losses.py, loaded verbatim through oracle/refshim.py) on seeded inputs, CPU,
fp32.  Run in the build container (needs /root/reference):

    python -m oracle.make_golden_masks

Writes tests/golden/ar_masks.pt.  Both branches exist for gridded (convolutional) models only:
on flattened graph inputs the reference's `_next_x` / `mask_tensor` raise (IndexError / ValueError).
"""

from __future__ import annotations

import os
import sys

import torch

from . import mfai_shim, refshim
from .make_golden import GOLDEN_DIR, NAMES5, LinearTanh, dataset_info_for, make_problem, pack_problem


def sprinkle_nans(p, seed):
    """NaNs in different pixels of every input state, of two forcing steps and of the truth."""
    g = torch.Generator().manual_seed(seed)
    s = p.shape

    def hit(t, frac):
        m = torch.rand(t.shape[:-1], generator=g) < frac  # whole pixels or single features
        one = torch.rand(t.shape, generator=g) < 0.5
        t[m.unsqueeze(-1) & one] = float("nan")

    hit(p.inputs, 0.06)
    hit(p.forcing, 0.04)
    hit(p.outputs, 0.08)
    # a pixel that is NaN in EVERY feature of the truth at every step and sample: masked-everywhere pixel
    p.outputs[:, :, s["H"] // 2, s["W"] // 2, :] = float("nan")


def run_case(p, strategy, mask_on_nan=False, mask_ratio=0.0, model_seed=7, rng_seed=None):
    ref_losses, _ = refshim.load_reference()
    NT = mfai_shim.NamedTensor
    s = p.shape
    ds = strategy == "downscaling_only"
    f_in = s["T_in"] * s["F"] * (1 - ds) + s["Fs"] + s["Ff"] + int(mask_on_nan)
    model = LinearTanh(f_in, s["F"], False, True, model_seed)
    info = dataset_info_for(p)
    lm = refshim.make_reference_module(
        model, training_strategy=strategy, border_mask=p.border_mask,
        grid_static_features=mfai_shim.expand_to_batch(p.statics, s["B"]),
        diff_stats=info.diff_stats, stats=info.stats, mask_on_nan=mask_on_nan, mask_ratio=mask_ratio,
    )
    batch = refshim.ItemBatch(
        inputs=NT(p.inputs.clone(), NAMES5, p.feature_names),
        forcing=NT(p.forcing.clone(), NAMES5, p.forcing_names),
        outputs=NT(p.outputs.clone(), NAMES5, p.feature_names),
    )
    out = {}
    if rng_seed is not None:
        torch.manual_seed(rng_seed)
        st = torch.get_rng_state()
        out["block_masks"] = []
        for _ in range(s["T"]):  # what mask_tensor paints for each model call, from ones
            out["block_masks"].append(lm.mask_tensor(torch.ones(1, s["H"], s["W"], 1))[0, :, :, 0].bool())
        torch.set_rng_state(st)
    pred, target = lm.common_step(batch, 0, "train")
    mask, target_masked = lm.get_mask_on_nan(target)  # lightning.py:813
    holder = torch.nn.Module()
    wl = ref_losses.WeightedLoss("MSELoss", reduction="none")
    wl.prepare(holder, 1.0 - p.border_mask, info)
    loss_bt = wl(pred, target_masked, mask)
    batch_loss = torch.mean(loss_bt)  # lightning.py:816
    (dW,) = torch.autograd.grad(batch_loss, model.W, retain_graph=True)
    sl = ref_losses.ScaledLoss("L1Loss", reduction="none")
    sl.prepare(torch.nn.Module(), 1.0 - p.border_mask, info)
    mae = sl(pred, target_masked, mask)
    out.update({
        "pred": pred.tensor.detach().clone(), "weighted_mse_bt": loss_bt.detach().clone(),
        "batch_loss": batch_loss.detach().clone(), "dW": dW.clone(), "scaled_mae_btf": mae.detach().clone(),
        "model_W": model.W.detach().clone(), "target_masked": target_masked.tensor.clone(),
        "mask": mask.bool().clone(),
    })
    return out


def main():
    if not refshim.reference_available():
        print("reference sources not available: cannot regenerate goldens", file=sys.stderr)
        return 1
    cases = {}
    for name, kw, shape in (
        ("nan_scaled_ar", dict(strategy="scaled_ar", mask_on_nan=True),
         dict(B=2, T_in=2, T=4, H=16, W=20, F=3, Ff=2, border=2)),
        ("nan_diff_ar", dict(strategy="diff_ar", mask_on_nan=True),
         dict(B=2, T_in=3, T=3, H=12, W=14, F=4, Ff=3, border=2)),
        ("nan_scaled_ar_tin1", dict(strategy="scaled_ar", mask_on_nan=True),
         dict(B=1, T_in=1, T=2, H=10, W=12, F=3, Ff=1, border=2)),
    ):
        p = make_problem(2234, **shape)
        sprinkle_nans(p, 77)
        cases[name] = {"problem": pack_problem(p), "config": kw, "expected": run_case(p, **kw)}
    for name, kw, shape in (
        ("mae_scaled_ar", dict(strategy="scaled_ar", mask_ratio=0.6, rng_seed=4242),
         dict(B=2, T_in=2, T=3, H=16, W=20, F=3, Ff=2, border=2)),
        ("mae_diff_ar_odd", dict(strategy="diff_ar", mask_ratio=0.35, rng_seed=4243),
         dict(B=1, T_in=1, T=2, H=13, W=18, F=3, Ff=2, border=2)),
        ("mae_and_nan", dict(strategy="scaled_ar", mask_ratio=0.8, mask_on_nan=True, rng_seed=4244),
         dict(B=2, T_in=2, T=2, H=12, W=12, F=3, Ff=2, border=2)),
    ):
        p = make_problem(2235, **shape)
        if kw.get("mask_on_nan"):
            sprinkle_nans(p, 78)
        cases[name] = {"problem": pack_problem(p), "config": kw, "expected": run_case(p, **kw)}
    # downscaling_only (lightning.py:546-558, 611-621, 759-765): outputs are matched to forcing features by name
    for name, kw, shape in (
        ("ds_plain", dict(strategy="downscaling_only"), dict(B=2, T_in=1, T=3, H=16, W=20, F=3, Ff=4, border=2)),
        ("ds_nan", dict(strategy="downscaling_only", mask_on_nan=True),
         dict(B=2, T_in=2, T=2, H=12, W=14, F=3, Ff=4, border=2)),
    ):
        p = make_problem(2236, **shape)
        p.feature_names = ["aro_t2m_2m", "aro_u10_10m", "aro_r2_2m"]
        p.forcing_names = ["arp_u10_10m", "arp_t2m_2m", "arp_r2_2m", "arp_tp_0m"]
        if kw.get("mask_on_nan"):
            sprinkle_nans(p, 79)
        cases[name] = {"problem": pack_problem(p), "config": kw, "expected": run_case(p, **kw)}
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, "ar_masks.pt")
    torch.save(cases, path)
    print(path, os.path.getsize(path), "bytes")
    for k, v in cases.items():
        e = v["expected"]
        print(k, "loss", float(e["batch_loss"]), "nan in pred:", bool(torch.isnan(e["pred"]).any()),
              "masked frac", 1 - float(e["mask"].float().mean()))
    return 0


if __name__ == "__main__":
    sys.exit(main())
