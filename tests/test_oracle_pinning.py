"""Pins oracle/ar_ref.py (the restatement that travels to the GPU box) against
(1) the committed golden vectors generated from the verbatim reference and
(2) the verbatim reference itself on fresh random inputs, where /root/reference exists."""
import pytest
import torch

from helpers import assert_close, load_golden
from oracle import ar_ref, refshim


def _linear_tanh(W, features_second):
    def f(x):
        if features_second:
            return (0.5 * torch.tanh(x.movedim(1, -1) @ W)).movedim(-1, 1)
        return 0.5 * torch.tanh(x @ W)

    return f


def run_oracle_case(case):
    p, cfg, exp = case["problem"], case["config"], case["expected"]
    s = p["shape"]
    graph = cfg["graph"]
    W = exp["model_W"].clone().requires_grad_(True)
    flat = (lambda t, k: t.flatten(k, k + 1)) if graph else (lambda t, k: t)
    pred = ar_ref.ar_rollout(
        _linear_tanh(W, cfg.get("features_second", False)),
        flat(p["inputs"], 2), flat(p["forcing"], 2), flat(p["outputs"], 2),
        flat(p["statics"], 0).unsqueeze(0).expand(s["B"], *flat(p["statics"], 0).shape),
        flat(p["border_mask"], 0), 1.0 - flat(p["border_mask"], 0),
        p["diff_std"], p["diff_mean"], cfg["strategy"], cfg.get("num_inter_steps", 1),
        features_second=cfg.get("features_second", False),
    )
    pred5 = pred.reshape(s["B"], s["T"], s["H"], s["W"], s["F"])
    interior = 1.0 - p["border_mask"]
    w = p["state_weight"] / p["diff_std"] ** 2
    mask = torch.ones_like(p["outputs"])
    loss_bt = ar_ref.weighted_loss(pred5, p["outputs"], mask, w, interior.squeeze(-1), interior.sum().item())
    scaled = ar_ref.scaled_loss(pred5, p["outputs"], mask, p["stat_std"], interior, interior.sum().item())
    (dW,) = torch.autograd.grad(loss_bt.mean(), W, retain_graph=True)
    (dWs,) = torch.autograd.grad(scaled.mean(), W)
    return pred5, loss_bt, scaled, dW, dWs


@pytest.mark.parametrize(
    "name",
    ["graph_scaled_ar", "graph_diff_ar", "conv_scaled_ar", "conv_diff_ar", "graph_scaled_ar_inter2",
     "graph_scaled_ar_tin3", "config1_conv_64x64"],
)
def test_ar_oracle_matches_reference_golden(name):
    case = load_golden("ar_rollout.pt")[name]
    exp = case["expected"]
    pred5, loss_bt, scaled, dW, dWs = run_oracle_case(case)
    assert_close(pred5, exp["pred"], 1e-6, "prediction")
    assert_close(loss_bt, exp["weighted_mse_bt"], 1e-6, "WeightedLoss")
    assert_close(scaled, exp["scaled_rmse_btf"], 1e-6, "ScaledLoss")
    assert_close(dW, exp["dW"], 1e-5, "dW (weighted)")
    assert_close(dWs, exp["dW_scaled"], 1e-5, "dW (scaled)")


def test_scaled_ar_border_is_truth_exactly():
    case = load_golden("ar_rollout.pt")["graph_scaled_ar"]
    p = case["problem"]
    border = p["border_mask"].squeeze(-1).bool()
    assert torch.equal(case["expected"]["pred"][:, :, border], p["outputs"][:, :, border])


def test_loss_oracle_matches_reference_golden():
    g = load_golden("losses.pt")
    p = g["problem"]
    interior = 1.0 - p["border_mask"]
    ni = interior.sum().item()
    for key, exp in g["cases"].items():
        kind, mname, form = key.split("/")
        mask = torch.ones_like(g["target"]) if mname == "full" else g["mask_partial"]
        pred = g["pred"].clone().requires_grad_(True)
        if form == "scaled":
            val = ar_ref.scaled_loss(pred, g["target"], mask, p["stat_std"], interior, ni, kind)
        else:
            w = p["state_weight"] / p["diff_std"] ** (2.0 if kind == "MSELoss" else 1.0)
            val = ar_ref.weighted_loss(pred, g["target"], mask, w, interior.squeeze(-1), ni, kind,
                                       reduce_spatial_dim=(form == "weighted"))
        (dpred,) = torch.autograd.grad((val * exp["gout"]).sum(), pred)
        assert_close(val, exp["value"], 1e-6, key)
        assert_close(dpred, exp["dpred"], 1e-6, key + " dpred")


def test_loss_4d_equals_5d_view():
    """The generalised union-mask dim (-1) on (B,T,N,F) equals the reference on (B,T,H,W,F)."""
    g = load_golden("losses.pt")
    p = g["problem"]
    interior = 1.0 - p["border_mask"]
    w = p["state_weight"] / p["diff_std"] ** 2
    exp = g["cases"]["MSELoss/partial/weighted"]["value"]
    val = ar_ref.weighted_loss(
        g["pred"].flatten(2, 3), g["target"].flatten(2, 3), g["mask_partial"].flatten(2, 3), w,
        interior.flatten(), interior.sum().item(),
    )
    assert_close(val, exp, 1e-6)


@pytest.mark.skipif(not refshim.reference_available(), reason="needs /root/reference")
def test_reference_raises_on_4d_graph_tensors():
    """Documents the reference bug the oracle works around (losses.py:156)."""
    from oracle import mfai_shim

    ref_losses, _ = refshim.load_reference()
    info = __import__("oracle.make_golden", fromlist=["x"]).dataset_info_for(
        __import__("oracle.make_golden", fromlist=["x"]).make_problem(1, 1, 1, 1, 6, 6, 2, 1)
    )
    L = ref_losses.WeightedLoss("MSELoss", reduction="none")
    L.prepare(torch.nn.Module(), torch.ones(36, 1), info)
    t = torch.randn(1, 1, 36, 2)
    nt = mfai_shim.NamedTensor(t, ["batch", "timestep", "ngrid", "features"], ["w0", "w1"])
    with pytest.raises(IndexError):
        L(nt, nt, torch.ones_like(t))


@pytest.mark.skipif(not refshim.reference_available(), reason="needs /root/reference")
@pytest.mark.parametrize("strategy", ["scaled_ar", "diff_ar"])
def test_ar_oracle_matches_verbatim_reference_fresh_inputs(strategy):
    from oracle import make_golden

    p = make_golden.make_problem(777, B=1, T_in=2, T=3, H=10, W=14, F=4, Ff=3, border=2)
    exp = make_golden.run_reference_case(p, strategy, graph=True, model_seed=3)
    case = {"problem": make_golden.pack_problem(p), "config": dict(strategy=strategy, graph=True), "expected": exp}
    pred5, loss_bt, scaled, dW, dWs = run_oracle_case(case)
    assert_close(pred5, exp["pred"], 1e-6)
    assert_close(dW, exp["dW"], 1e-5)
