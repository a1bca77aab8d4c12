"""Whole-model parity (GraphLAM / HiLAM / HiLAMParallel) against the oracle restatement with the
same weights, and the reference's own model test (tests/test_models.py:63-142) mirrored."""
import pytest
import torch

from helpers import BF16_TOL, FP32_TOL, REPORT, assert_close, pixel_meshgrid

pytestmark = pytest.mark.gpu

ORACLES = {"GraphLAM": "GraphLAM", "HiLAM": "HiLAM", "HiLAMParallel": "HiLAMParallel"}


def _build(name, H, W, f_in, f_out, tmp_path, meshgrid=None, **settings):
    from oracle import graph_ref
    from py4cast_b200 import models
    from py4cast_b200.graph_build import build_mesh_graph

    kls, s = models.get_model_kls_and_settings(name, {"tmp_dir": str(tmp_path), **settings})
    mg = meshgrid if meshgrid is not None else pixel_meshgrid(H, W)
    kls.rank_zero_setup(s, mg)
    torch.manual_seed(0)
    model = kls(in_channels=f_in, out_channels=f_out, settings=s, input_shape=(H, W))
    graph = build_mesh_graph(mg, kls.hierarchical, s.n_max_levels)
    ref = getattr(graph_ref, ORACLES[name])(
        f_in, f_out, graph, processor_layers=s.processor_layers, mesh_aggr=s.mesh_aggr
    )
    ref.load_state_dict(model.state_dict())
    return model, ref


@pytest.mark.parametrize("name", ["GraphLAM", "HiLAM", "HiLAMParallel"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_model_forward_backward_vs_oracle(cuda, tmp_path, name, dtype):
    H, W, f_in, f_out, B = 64, 64, 11, 5, 2
    model, ref = _build(name, H, W, f_in, f_out, tmp_path, processor_layers=2)
    model = model.to(cuda)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, H * W, f_in, generator=g)
    gy = torch.randn(B, H * W, f_out, generator=g)
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    yr.backward(gy)
    xg = x.to(cuda).to(dtype).requires_grad_(True)
    y = model(xg)
    assert y.dtype == dtype and y.shape == (B, H * W, f_out)
    y.backward(gy.to(cuda).to(dtype))
    tol = FP32_TOL if dtype == torch.float32 else BF16_TOL
    assert_close(y.float().cpu(), yr, tol, f"{name} y")
    assert_close(xg.grad.float().cpu(), xr.grad, tol, f"{name} dx")
    flat_g = torch.cat([p.grad.flatten().cpu() for p in model.parameters()])
    flat_r = torch.cat([p.grad.flatten() for p in ref.parameters()])
    assert_close(flat_g, flat_r, tol, f"{name} all parameter grads")
    worst, worst_name = 0.0, ""
    for (n, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
        e = ((p.grad.float().cpu() - q.grad).norm() / (q.grad.norm() + 1e-30)).item()
        if e > worst:
            worst, worst_name = e, n
    REPORT.append((f"{name} {dtype}", f"worst per-parameter gradient error ({worst_name})", worst, 5 * tol))
    assert worst <= 5 * tol, f"worst per-parameter gradient error {worst:.2e} ({worst_name})"


def test_model_under_bf16_autocast_selects_bf16(cuda, tmp_path):
    model, ref = _build("GraphLAM", 64, 64, 6, 3, tmp_path, processor_layers=1)
    model = model.to(cuda)
    x = torch.randn(1, 4096, 6)
    with torch.amp.autocast("cuda", dtype=torch.bfloat16):
        y = model(x.to(cuda))
    assert y.dtype == torch.bfloat16
    assert_close(y.float().cpu(), ref(x), BF16_TOL)


def test_mean_aggregation_and_non_square_grid(cuda, tmp_path):
    model, ref = _build("GraphLAM", 40, 64, 4, 2, tmp_path, processor_layers=1, mesh_aggr="mean")
    model = model.to(cuda)
    x = torch.randn(1, 40 * 64, 4)
    assert_close(model(x.to(cuda)).cpu(), ref(x), FP32_TOL)


def test_forward_is_bit_reproducible(cuda, tmp_path):
    model, _ = _build("HiLAM", 64, 64, 8, 4, tmp_path, processor_layers=1)
    model = model.to(cuda)
    x = torch.randn(2, 4096, 8, device=cuda)
    grads = []
    for _ in range(2):
        model.zero_grad(set_to_none=True)
        y = model(x)
        y.square().mean().backward()
        grads.append((y.detach().clone(), torch.cat([p.grad.flatten() for p in model.parameters()])))
    assert torch.equal(grads[0][0], grads[1][0]) and torch.equal(grads[0][1], grads[1][1])


@pytest.mark.parametrize("model_name", ["GraphLAM", "HiLAM", "HiLAMParallel", "HalfUNet"])
def test_torch_training_loop(cuda, tmp_path, model_name):
    """Reference tests/test_models.py:63-142: instantiate with default settings through the
    registry, 2 SGD steps of MSELoss on y = sum(x), eval-mode forward."""
    from py4cast_b200.base import ModelType, features_last_to_second, features_second_to_last
    from py4cast_b200.models import get_model_kls_and_settings

    H = W = 64
    settings = {"tmp_dir": str(tmp_path)} if model_name != "HalfUNet" else {}
    model_kls, model_settings = get_model_kls_and_settings(model_name, settings)
    if hasattr(model_kls, "rank_zero_setup"):
        model_kls.rank_zero_setup(model_settings, pixel_meshgrid(H, W))
    model = model_kls(in_channels=2, out_channels=1, settings=model_settings, input_shape=(H, W)).to(cuda)
    optimizer = torch.optim.SGD(model.parameters(), lr=0.001, momentum=0.9)
    loss_fn = torch.nn.MSELoss()
    losses = []
    for _ in range(2):
        inputs = torch.rand(2, H, W, 2, device=cuda)
        targets = inputs.sum(-1, keepdim=True)
        optimizer.zero_grad()
        if model.model_type == ModelType.GRAPH:
            inputs, targets = inputs.flatten(1, 2), targets.flatten(1, 2)
        if model.features_second:
            outputs = features_second_to_last(model(features_last_to_second(inputs)))
        else:
            outputs = model(inputs)
        loss = loss_fn(outputs, targets)
        loss.backward()
        optimizer.step()
        losses.append(loss.item())
    assert all(torch.isfinite(torch.tensor(losses)))
    model.eval()
    sample = torch.rand(1, H, W, 2, device=cuda)
    sample = sample.flatten(1, 2) if model.model_type == ModelType.GRAPH else features_last_to_second(sample)
    model(sample)


def test_titan_hilam_one_step_parity_fp32(cuda, tmp_path):
    """BASELINE-size graph (512x640, 3 levels, 1.3 M m2g edges): forward + backward vs the oracle."""
    from py4cast_b200.graph_build import unit_meshgrid

    H, W = 512, 640
    torch.set_num_threads(max(1, torch.get_num_threads()))
    model, ref = _build("HiLAM", H, W, 67, 21, tmp_path, meshgrid=unit_meshgrid(H, W), n_max_levels=3)
    model = model.to(cuda)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(1, H * W, 67, generator=g)
    gy = torch.randn(1, H * W, 21, generator=g)
    yr = ref(x)
    yr.backward(gy)
    y = model(x.to(cuda))
    y.backward(gy.to(cuda))
    assert_close(y.cpu(), yr, FP32_TOL, "titan HiLAM y")
    flat_g = torch.cat([p.grad.flatten().cpu() for p in model.parameters()])
    flat_r = torch.cat([p.grad.flatten() for p in ref.parameters()])
    assert_close(flat_g, flat_r, FP32_TOL, "titan HiLAM parameter grads")
