"""BASELINE configs[4] (and [0]): a `features_second` convolutional model pushed through OUR AR-step and loss
kernels.  The CNN itself is the stock-cuDNN stand-in of py4cast_b200/standins.py (the reference's UNet /
Segformer live in the absent mfai package and are out of scope as kernels), so the line reports the whole
training step AND the share of it spent in this repo's kernels (AR concat / update / blend, loss forward /
backward, AdamW), measured by timing the same step with the model replaced by a zero-cost stub.
usage: conv_path_bench.py [--grid 1024x1024] [--batch 2] [--pred-steps 4] [--precision bf16-true]"""
import argparse
import json
import os
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class _Stub(torch.nn.Module):
    """y = w * x[:, :F_out]: keeps autograd and the layout contract, costs one element-wise kernel."""

    def __init__(self, f_out):
        super().__init__()
        self.w = torch.nn.Parameter(torch.ones(()))
        self.f_out = f_out

    def forward(self, x):
        return x[:, : self.f_out] * self.w


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", default="1024x1024")
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--pred-steps", type=int, default=4)
    ap.add_argument("--precision", default="bf16-true")
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    from py4cast_b200.datasets import synthetic_batch, synthetic_dataset_info
    from py4cast_b200.lightning import AutoRegressiveLightning
    from py4cast_b200.trainer import DataParallelTrainer

    dev = torch.device("cuda:0")
    H, W = (int(v) for v in a.grid.split("x"))
    info = synthetic_dataset_info(H, W)
    res = {}
    for label in ("HalfUNet stand-in", "stub model (our kernels only)"):
        torch.manual_seed(0)
        lm = AutoRegressiveLightning(
            settings_init_args={}, dataset_info=info, num_input_steps=2, num_pred_steps_train=a.pred_steps,
            batch_size=a.batch, model_name="HalfUNet",
            losses=[{"class": "WeightedLoss", "weight": 1.0, "params": {"loss": "MSELoss", "reduction": "none"}}],
            training_strategy="scaled_ar", precision=a.precision,
        )
        if label.startswith("stub"):
            stub = _Stub(info.weather_dim)
            for attr in ("model_type", "features_last", "features_second", "num_spatial_dims"):
                setattr(_Stub, attr, getattr(type(lm.model), attr, getattr(lm.model, attr, None)))
            lm.model = stub
        lm = lm.to(dev)
        tr = DataParallelTrainer(lm, total_steps=1000)
        batch = synthetic_batch(info, a.batch, 2, a.pred_steps, seed=0, device=dev)
        for _ in range(3):
            tr.train_step(batch)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            loss = tr.train_step(batch)
        e1.record()
        torch.cuda.synchronize()
        res[label] = e0.elapsed_time(e1) / a.steps
        del tr, lm
        torch.cuda.empty_cache()
    full, ours = res["HalfUNet stand-in"], res["stub model (our kernels only)"]
    print(json.dumps({
        "config": f"conv stand-in {H}x{W}, scaled_ar T={a.pred_steps}, WeightedLoss(MSE), {a.precision}, batch {a.batch}",
        "ms_per_step": full, "samples_per_s": a.batch / (full * 1e-3),
        "ms_per_step_ar_loss_optimizer_kernels_only": ours, "share_of_step_in_this_repos_kernels": ours / full,
        "final_loss": float(loss)}))


if __name__ == "__main__":
    main()
