"""AR rollout + losses on the CUDA kernels vs the golden vectors produced by the VERBATIM
reference (tests/golden/*.pt, oracle/make_golden.py) — through the reference-facing API:
AutoRegressiveLightning.common_step / training_step and WeightedLoss / ScaledLoss / CombinedLoss."""
import datetime as dt
from dataclasses import dataclass

import pytest
import torch
from torch import nn

from helpers import FP32_TOL, assert_close, load_golden

pytestmark = pytest.mark.gpu

NAMES5 = ["batch", "timestep", "lat", "lon", "features"]


def _dataset_info(p):
    from py4cast_b200.datasets import DatasetInfo, Statics, Stats
    from py4cast_b200.namedtensor import NamedTensor

    s = p["shape"]
    fn = p["feature_names"]
    statics = Statics(
        grid_statics=NamedTensor(p["statics"].clone(), ["lat", "lon", "features"],
                                 ["x", "y", "geopotential", "border_mask"][: s["Fs"]]),
        grid_shape=(s["H"], s["W"]),
    )
    stats = Stats(stats={n: {"std": p["stat_std"][i], "mean": torch.tensor(0.0)} for i, n in enumerate(fn)})
    diff = Stats(stats={n: {"std": p["diff_std"][i], "mean": p["diff_mean"][i]} for i, n in enumerate(fn)})
    return DatasetInfo(
        name="golden", domain_info=None, units={}, weather_dim=s["F"], forcing_dim=s["Ff"],
        pred_step=dt.timedelta(hours=1), statics=statics, stats=stats, diff_stats=diff,
        state_weights={n: float(p["state_weight"][i]) for i, n in enumerate(fn)},
    )


def _register_linear_tanh():
    """Plugin models implementing the golden generator's stand-in network y = 0.5 tanh(x W)."""
    from py4cast_b200 import models
    from py4cast_b200.base import ModelABC, ModelType

    @dataclass(slots=True)
    class S:
        pass

    def make(name, graph, second):
        class M(ModelABC, nn.Module):
            settings_kls = S
            onnx_supported = False
            supported_num_spatial_dims = (1,) if graph else (2,)
            num_spatial_dims = 1 if graph else 2
            features_last = not second
            model_type = ModelType.GRAPH if graph else ModelType.CONVOLUTIONAL
            register = True

            def __init__(self, in_channels, out_channels, input_shape, settings, *a, **k):
                super().__init__()
                self.in_channels, self.out_channels, self.input_shape = in_channels, out_channels, input_shape
                self.W = nn.Parameter(torch.zeros(in_channels, out_channels))

            @property
            def settings(self):
                return S()

            def forward(self, x):
                if second:
                    return (0.5 * torch.tanh(x.movedim(1, -1) @ self.W)).movedim(-1, 1)
                return 0.5 * torch.tanh(x @ self.W)

        M.__name__ = name
        models.registry[name] = M

    make("GoldenGraphNet", True, False)
    make("GoldenConvNet", False, True)


def _batch(p, dev):
    from py4cast_b200.datasets import ItemBatch
    from py4cast_b200.namedtensor import NamedTensor

    return ItemBatch(
        inputs=NamedTensor(p["inputs"].clone().to(dev), NAMES5, p["feature_names"]),
        forcing=NamedTensor(p["forcing"].clone().to(dev), NAMES5, p["forcing_names"]),
        outputs=NamedTensor(p["outputs"].clone().to(dev), NAMES5, p["feature_names"]),
    )


@pytest.mark.parametrize(
    "name",
    ["graph_scaled_ar", "graph_diff_ar", "conv_scaled_ar", "conv_diff_ar", "graph_scaled_ar_inter2",
     "graph_scaled_ar_tin3", "config1_conv_64x64"],
)
def test_ar_rollout_and_loss_vs_reference_golden(cuda, name):
    from py4cast_b200.lightning import AutoRegressiveLightning
    from py4cast_b200.losses import ScaledLoss

    _register_linear_tanh()
    case = load_golden("ar_rollout.pt")[name]
    p, cfg, exp = case["problem"], case["config"], case["expected"]
    s = p["shape"]
    info = _dataset_info(p)
    lm = AutoRegressiveLightning(
        settings_init_args={}, dataset_info=info, num_input_steps=s["T_in"],
        num_pred_steps_train=s["T"], batch_size=s["B"],
        model_name="GoldenGraphNet" if cfg["graph"] else "GoldenConvNet",
        losses=[{"class": "WeightedLoss", "weight": 1.0, "params": {"loss": "MSELoss", "reduction": "none"}}],
        training_strategy=cfg["strategy"], num_inter_steps=cfg.get("num_inter_steps", 1),
    )
    with torch.no_grad():
        lm.model.W.copy_(exp["model_W"])
    lm = lm.to(cuda)

    batch = _batch(p, cuda)
    pred, target = lm.common_step(batch, 0, "train")
    expected_names = ["batch", "timestep", "ngrid", "features"] if cfg["graph"] else NAMES5
    assert pred.names == expected_names and target is batch.outputs
    pred5 = pred.tensor.reshape(s["B"], s["T"], s["H"], s["W"], s["F"])
    assert_close(pred5.cpu(), exp["pred"], FP32_TOL, f"{name} prediction")
    if cfg["strategy"] == "scaled_ar":  # border forced to the truth, exactly
        border = p["border_mask"].squeeze(-1).bool()
        assert torch.equal(pred5.cpu()[:, :, border], p["outputs"][:, :, border])

    loss_bt = lm.loss(pred, target, mask=None)
    assert loss_bt.shape == (s["B"], s["T"])
    assert_close(loss_bt.cpu(), exp["weighted_mse_bt"], FP32_TOL, f"{name} WeightedLoss")
    batch_loss = torch.mean(loss_bt)
    (dW,) = torch.autograd.grad(batch_loss, lm.model.W, retain_graph=True)
    assert_close(dW.cpu(), exp["dW"], FP32_TOL, f"{name} d(batch_loss)/dW")

    sl = ScaledLoss("MSELoss", reduction="none")
    sl.prepare(lm, lm.interior_mask, info)
    scaled = sl(pred, target, mask=None)
    assert scaled.shape == (s["B"], s["T"], s["F"])
    assert_close(scaled.cpu(), exp["scaled_rmse_btf"], FP32_TOL, f"{name} ScaledLoss")
    (dWs,) = torch.autograd.grad(scaled.mean(), lm.model.W)
    assert_close(dWs.cpu(), exp["dW_scaled"], FP32_TOL, f"{name} d(scaled)/dW")

    # training_step == mean(loss) on a fresh batch (common_step mutates the batch in place)
    tl = lm.training_step(_batch(p, cuda), 1)
    assert_close(tl.detach().cpu(), exp["batch_loss"], FP32_TOL, f"{name} training_step")


def test_inference_phase_has_no_border_forcing(cuda):
    from py4cast_b200.lightning import AutoRegressiveLightning

    _register_linear_tanh()
    case = load_golden("ar_rollout.pt")["graph_scaled_ar"]
    p, exp = case["problem"], case["expected"]
    s = p["shape"]
    lm = AutoRegressiveLightning({}, _dataset_info(p), num_input_steps=2, batch_size=2,
                                 model_name="GoldenGraphNet", training_strategy="scaled_ar")
    with torch.no_grad():
        lm.model.W.copy_(exp["model_W"])
    lm = lm.to(cuda)
    lm.common_step(_batch(p, cuda), 0, "train")  # stores feature / dim names
    from oracle import ar_ref

    W = exp["model_W"]
    ref = ar_ref.ar_rollout(
        lambda x: 0.5 * torch.tanh(x @ W), p["inputs"].flatten(2, 3), p["forcing"].flatten(2, 3),
        p["outputs"].flatten(2, 3), p["statics"].flatten(0, 1).unsqueeze(0).expand(2, -1, -1),
        p["border_mask"].flatten(0, 1), 1 - p["border_mask"].flatten(0, 1), p["diff_std"], p["diff_mean"],
        "scaled_ar", phase="inference",
    )
    with torch.no_grad():
        out = lm(_batch(p, cuda), 1)
    assert_close(out.tensor.cpu(), ref, FP32_TOL, "inference rollout")


def test_losses_vs_reference_golden(cuda):
    from py4cast_b200.losses import CombinedLoss, ScaledLoss, WeightedLoss
    from py4cast_b200.namedtensor import NamedTensor

    g = load_golden("losses.pt")
    p = g["problem"]
    info = _dataset_info(p)
    interior = (1.0 - p["border_mask"]).to(cuda)
    for layout in ("5d", "4d"):
        for key, exp in g["cases"].items():
            kind, mname, form = key.split("/")
            if layout == "4d" and form == "weighted_map":
                continue
            flat = (lambda t: t.flatten(2, 3)) if layout == "4d" else (lambda t: t)
            names = ["batch", "timestep", "ngrid", "features"] if layout == "4d" else NAMES5
            pred = flat(g["pred"]).clone().to(cuda).requires_grad_(True)
            P = NamedTensor(pred, names, p["feature_names"])
            T = NamedTensor(flat(g["target"]).to(cuda), names, p["feature_names"])
            mask = None if mname == "full" else flat(g["mask_partial"]).to(cuda)
            holder = nn.Module()
            im = interior.flatten(0, 1) if layout == "4d" else interior
            if form == "scaled":
                L = ScaledLoss(kind, reduction="none")
                L.prepare(holder, im, info)
                val = L(P, T, mask)
            else:
                L = WeightedLoss(kind, reduction="none")
                L.prepare(holder, im, info)
                val = L(P, T, mask, reduce_spatial_dim=(form == "weighted"))
            assert val.shape == exp["value"].shape
            (dpred,) = torch.autograd.grad((val * exp["gout"].to(cuda)).sum(), pred)
            assert_close(val.cpu(), exp["value"], FP32_TOL, f"{layout}/{key}")
            assert_close(dpred.cpu().reshape(exp["dpred"].shape), exp["dpred"], FP32_TOL, f"{layout}/{key} dpred")
            if mname == "full":  # an explicit all-ones mask is the same as mask=None
                val1 = L(P, T, torch.ones_like(pred)) if form == "scaled" else L(
                    P, T, torch.ones_like(pred), reduce_spatial_dim=(form == "weighted"))
                assert torch.equal(val1, val)
    exp = g["combined"]
    pred = g["pred"].clone().to(cuda).requires_grad_(True)
    cl = CombinedLoss(exp["config"])
    cl.prepare(nn.Module(), interior, info)
    val = cl(NamedTensor(pred, NAMES5, p["feature_names"]), NamedTensor(g["target"].to(cuda), NAMES5, p["feature_names"]), None)
    (dpred,) = torch.autograd.grad(val.mean(), pred)
    assert_close(val.cpu(), exp["value"], FP32_TOL, "CombinedLoss")
    assert_close(dpred.cpu(), exp["dpred"], FP32_TOL, "CombinedLoss dpred")


def test_loss_properties_at_titan_size(cuda):
    """Size-independent properties at the BASELINE grid (512x640, F=21, B*T=4)."""
    from py4cast_b200.datasets import synthetic_dataset_info
    from py4cast_b200.losses import ScaledLoss, WeightedLoss
    from py4cast_b200.namedtensor import NamedTensor

    info = synthetic_dataset_info()
    names = ["batch", "timestep", "ngrid", "features"]
    fn = info.shortnames["input_output"]
    interior = info.statics.interior_mask.flatten(0, 1).to(cuda)
    holder = nn.Module()
    wl = WeightedLoss("MSELoss", reduction="none")
    wl.prepare(holder, interior, info)
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn(1, 4, 512 * 640, 21, device=cuda, generator=g)
    b = torch.randn(1, 4, 512 * 640, 21, device=cuda, generator=g)
    A, Bn = NamedTensor(a, names, fn), NamedTensor(b, names, fn)
    assert float(wl(A, A).abs().max()) == 0.0  # identical fields -> zero loss
    l1 = wl(A, Bn)
    assert torch.equal(l1, wl(A, Bn))  # bit-reproducible
    assert_close(wl(NamedTensor(2 * a, names, fn), NamedTensor(2 * b, names, fn)), 4 * l1, 1e-6, "MSE is 2-homogeneous")
    # against the torch fp32 formula on the GPU
    w = wl.weights(tuple(fn), cuda)
    ref = (((a - b) ** 2 * w).sum(-1) * interior.squeeze(-1)).sum(-1) / interior.sum()
    assert_close(l1, ref, FP32_TOL, "titan WeightedLoss")
    sl = ScaledLoss("L1Loss", reduction="none")
    sl.prepare(nn.Module(), interior, info)
    ref = ((a - b).abs() * interior).sum(2) / interior.sum() * sl.weights(tuple(fn), cuda)
    assert_close(sl(A, Bn), ref, FP32_TOL, "titan ScaledLoss(L1)")


@pytest.mark.parametrize("name", ["graph_scaled_ar", "conv_diff_ar"])
def test_validation_and_test_steps_vs_reference_golden(cuda, name):
    """validation_step / test_step (reference lightning.py:887-917, 1017-1042 + the device-side
    work of StateErrorPlot / SpatialErrorPlot): same rollout under no_grad, per-timestep loss,
    ScaledLoss MAE / RMSE per feature, spatial loss map — against the verbatim-reference golden
    vectors and the loss oracle."""
    from oracle import ar_ref
    from py4cast_b200.lightning import AutoRegressiveLightning

    _register_linear_tanh()
    case = load_golden("ar_rollout.pt")[name]
    p, cfg, exp = case["problem"], case["config"], case["expected"]
    s = p["shape"]
    info = _dataset_info(p)
    lm = AutoRegressiveLightning(
        settings_init_args={}, dataset_info=info, num_input_steps=s["T_in"], num_pred_steps_train=s["T"],
        num_pred_steps_val_test=s["T"], batch_size=s["B"],
        model_name="GoldenGraphNet" if cfg["graph"] else "GoldenConvNet",
        losses=[{"class": "WeightedLoss", "weight": 1.0, "params": {"loss": "MSELoss", "reduction": "none"}}],
        training_strategy=cfg["strategy"],
    )
    with torch.no_grad():
        lm.model.W.copy_(exp["model_W"])
    lm = lm.to(cuda)

    val = lm.validation_step(_batch(p, cuda), 0)
    assert not val["prediction"].tensor.requires_grad
    assert_close(val["time_step_loss"].cpu(), exp["weighted_mse_bt"].mean(0), FP32_TOL, f"{name} val time_step_loss")
    assert_close(val["mean_loss"].cpu(), exp["batch_loss"], FP32_TOL, f"{name} val mean_loss")
    assert lm.val_mean_loss is val["mean_loss"] and len(lm.validation_step_losses) == 1

    tst = lm.test_step(_batch(p, cuda), 0)
    assert_close(tst["rmse"].cpu(), exp["scaled_rmse_btf"], FP32_TOL, f"{name} test rmse")
    # MAE and the spatial map: loss oracle (pinned against the verbatim reference in
    # tests/test_oracle_pinning.py) on the golden prediction
    pred, tgt = exp["pred"].flatten(2, 3), p["outputs"].flatten(2, 3)
    interior1 = (1 - p["border_mask"]).flatten(0, 1)  # (N, 1)
    interior = interior1.squeeze(-1)
    ones = torch.ones_like(pred)
    mae = ar_ref.scaled_loss(pred, tgt, ones, p["stat_std"], interior1, interior.sum().item(), kind="L1Loss")
    assert tst["mae"].shape == (s["B"], s["T"], s["F"])
    assert_close(tst["mae"].cpu(), mae, FP32_TOL, f"{name} test mae")
    assert_close(val["mae"].cpu(), mae, FP32_TOL, f"{name} val mae")
    w = p["state_weight"] / p["diff_std"] ** 2
    smap = ar_ref.weighted_loss(pred, tgt, ones, w, interior, interior.sum().item(), reduce_spatial_dim=False)
    assert tst["spatial_loss"].shape == (s["B"], s["T"], s["H"], s["W"])
    assert_close(tst["spatial_loss"].cpu().flatten(2, 3), smap, FP32_TOL, f"{name} spatial loss map")


@pytest.mark.parametrize("T_in,x_dtype,y_dtype,scaled,border", [
    (2, torch.bfloat16, torch.bfloat16, False, False),   # the benchmarked configuration (diff_ar, bf16-true)
    (2, torch.bfloat16, torch.bfloat16, True, True),     # scaled_ar with the border blend
    (3, torch.float32, torch.float32, True, True),       # three input steps: the middle d_prev comes from g_xnext
    (1, torch.bfloat16, torch.float32, False, True),
])
def test_staged_ar_kernels_equal_the_scalar_ones_bit_for_bit(cuda, T_in, x_dtype, y_dtype, scaled, border):
    """The cp.async-staged concat / step kernel and the vectorised step backward (nodes % 64 == 0, aligned arrays)
    against the scalar kernels they replace (p4c_set_ar_vector_paths(0)): every output bit-identical; and against
    plain torch (reference lightning.py:598-633, 711-767) within fp32 rounding."""
    from py4cast_b200 import _lib
    from py4cast_b200 import lightning as L

    lib = _lib.load()
    B, N, F, Fs, Ff = 3, 64 * 7, 21, 4, 21
    g = torch.Generator().manual_seed(T_in * 7 + int(scaled))
    prevs = [torch.randn(B, N, F, generator=g).to(cuda) for _ in range(T_in)]
    static = torch.randn(N, Fs, generator=g).to(cuda)
    forcing0 = torch.randn(B, N, Ff, generator=g).to(cuda)
    forcing1 = torch.randn(B, N, Ff, generator=g).to(cuda)
    truth = torch.randn(B, N, F, generator=g).to(cuda)
    bmask = (torch.rand(N, generator=g) < 0.3).float().to(cuda)
    std = (torch.rand(F, generator=g) + 0.5).to(cuda)
    mean = torch.randn(F, generator=g).to(cuda)
    yv = torch.randn(B, N, F, generator=g).to(cuda).to(y_dtype)
    g_state = torch.randn(B, N, F, generator=g).to(cuda)

    def run():
        x0 = L.ar_concat(prevs, static, forcing0, x_dtype)
        ps = [p.clone().requires_grad_(True) for p in prevs]
        y = yv.clone().requires_grad_(True)
        cfg = {"truth": truth if border else None, "border": bmask if border else None,
               "std": std if scaled else None, "mean": mean if scaled else None,
               "new_state": torch.empty(B, N, F, device=cuda), "forcing_next": forcing1, "static": static,
               "x_dtype": x_dtype}
        new_state, x1 = L._ARStepFn.apply(cfg, y, *ps)
        g_x = torch.randn(x1.shape, generator=torch.Generator().manual_seed(1)).to(cuda).to(x1.dtype)
        torch.autograd.backward([new_state, x1], [g_state, g_x])
        outs = [x0, new_state, x1, y.grad] + [p.grad for p in ps if p.grad is not None]
        return [o.detach().clone() for o in outs]

    try:
        lib.p4c_set_ar_vector_paths(0)
        scalar = run()
        lib.p4c_set_ar_vector_paths(1)
        staged = run()
    finally:
        lib.p4c_set_ar_vector_paths(1)
    assert len(scalar) == len(staged)
    for i, (a, b) in enumerate(zip(scalar, staged)):
        assert a.shape == b.shape and a.stride() == b.stride()
        assert torch.equal(a, b), f"output {i}: staged / vectorised kernel differs from the scalar kernel"
    # plain-torch reference of the forward
    pred = prevs[-1] + (yv.float() * std + mean if scaled else yv.float())
    if border:
        pred = bmask[None, :, None] * truth + (1 - bmask[None, :, None]) * pred
    assert_close(staged[1], pred, FP32_TOL, "new state vs torch")
    x1_ref = torch.cat(prevs[1:] + [pred, static.expand(B, -1, -1), forcing1], dim=-1)
    f_in = x1_ref.shape[-1]
    assert_close(staged[2][..., :f_in].float(), x1_ref.to(x_dtype).float(), 1e-2 if x_dtype == torch.bfloat16 else FP32_TOL,
                 "x_next vs torch")
    assert torch.equal(staged[0][..., :f_in],
                       torch.cat(prevs + [static.expand(B, -1, -1), forcing0], dim=-1).to(x_dtype)), "concat is a pure copy"
