"""Training step on the GPU: fused AdamW + device-side LR schedule vs torch.optim.AdamW +
transformers' scheduler, gradient sinks vs autograd-returned gradients, CUDA-graph replay vs
eager, and a full small HiLAM step against the oracle pipeline."""
import ctypes

import pytest
import torch

from helpers import FP32_TOL, assert_close, pixel_meshgrid

pytestmark = pytest.mark.gpu


def test_fused_adamw_matches_torch_adamw_and_schedule(cuda):
    from transformers.optimization import get_cosine_with_min_lr_schedule_with_warmup

    from py4cast_b200.trainer import FlatParams, FusedAdamW

    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(33, 17), torch.nn.Linear(17, 5)).to(cuda)
    ref = torch.nn.Sequential(torch.nn.Linear(33, 17), torch.nn.Linear(17, 5)).to(cuda)
    ref.load_state_dict(net.state_dict())
    base, min_lr, warm, total, betas = 1e-3, 3e-7, 3, 40, (0.9, 0.95)
    fp = FlatParams(net)
    opt = FusedAdamW(fp, base, min_lr, warm, total, betas=betas)
    topt = torch.optim.AdamW(ref.parameters(), lr=base, betas=betas)
    sched = get_cosine_with_min_lr_schedule_with_warmup(topt, num_warmup_steps=warm, num_training_steps=total, min_lr=min_lr)
    g = torch.Generator(device="cuda").manual_seed(1)
    for step in range(12):
        x = torch.randn(16, 33, device=cuda, generator=g)
        fp.zero_grad()
        net(x).square().mean().backward()
        lr_expected = sched.get_last_lr()[0]
        opt.step()
        topt.zero_grad()
        ref(x).square().mean().backward()
        topt.step()
        sched.step()
        st = opt.host_state()
        assert st["step"] == step + 1 and abs(st["lr"] - lr_expected) <= 1e-9
        for p, q in zip(net.parameters(), ref.parameters()):
            assert_close(p, q, 1e-6, f"AdamW step {step}")


def _lm(tmp_path, name="HiLAM", H=64, W=64, strategy="scaled_ar", precision="32-true", T_in=2, pl=1):
    from py4cast_b200.datasets import synthetic_dataset_info
    from py4cast_b200.lightning import AutoRegressiveLightning

    info = synthetic_dataset_info(H, W, num_features=5, num_forcing=3, border_size=4)
    torch.manual_seed(0)
    lm = AutoRegressiveLightning(
        settings_init_args={"tmp_dir": str(tmp_path), "processor_layers": pl}, dataset_info=info,
        num_input_steps=T_in, num_pred_steps_train=3, batch_size=2, model_name=name,
        training_strategy=strategy, learning_rate=1e-3, min_learning_rate=1e-6, betas=(0.9, 0.95),
        precision=precision,
    )
    return lm, info


def test_kernel_grad_sinks_equal_autograd_grads(cuda, tmp_path):
    from py4cast_b200.datasets import synthetic_batch
    from py4cast_b200.trainer import FlatParams, _fresh

    lm, info = _lm(tmp_path)
    lm = lm.to(cuda)
    batch = synthetic_batch(info, 2, 2, 3, device=cuda)
    lm.training_step(_fresh(batch), 1).backward()
    plain = torch.cat([p.grad.flatten() for p in lm.parameters()])
    lm.zero_grad(set_to_none=True)
    fp = FlatParams(lm)
    fp.enable_kernel_sinks(True)
    lm.training_step(_fresh(batch), 1).backward()
    sunk = torch.cat([p.grad.flatten() for p in lm.parameters()])
    assert torch.equal(plain, sunk)


@pytest.mark.parametrize("name,strategy", [("HiLAM", "scaled_ar"), ("GraphLAM", "diff_ar")])
def test_full_training_steps_vs_oracle_pipeline(cuda, tmp_path, name, strategy):
    """2 optimiser steps: CUDA path (kernels + fused AdamW) vs oracle model + oracle AR/loss
    restatement + torch AdamW on CPU, same weights and data."""
    from oracle import ar_ref, graph_ref
    from py4cast_b200.datasets import synthetic_batch
    from py4cast_b200.graph_build import build_mesh_graph
    from py4cast_b200.trainer import DataParallelTrainer

    lm, info = _lm(tmp_path, name=name, strategy=strategy)
    H, W = info.statics.grid_shape
    graph = build_mesh_graph(info.statics.meshgrid, lm.model.hierarchical)
    ref = getattr(graph_ref, name)(lm.model.in_channels, lm.model.out_channels, graph, processor_layers=1)
    ref.load_state_dict(lm.model.state_dict())
    lm = lm.to(cuda)
    tr = DataParallelTrainer(lm, total_steps=100)
    topt = torch.optim.AdamW(ref.parameters(), lr=1e-3, betas=(0.9, 0.95))
    fn = info.shortnames["input_output"]
    std, mean = info.diff_stats.to_list("std", fn), info.diff_stats.to_list("mean", fn)
    w = torch.tensor([info.state_weights[n] for n in fn]) / std**2
    border = info.statics.border_mask.flatten(0, 1)
    static = info.statics.grid_statics.tensor.flatten(0, 1)
    for step in range(2):
        batch = synthetic_batch(info, 2, 2, 3, seed=step, device="cpu")
        loss = tr.train_step(synthetic_batch(info, 2, 2, 3, seed=step, device=cuda))
        pred = ar_ref.ar_rollout(
            ref, batch.inputs.tensor.flatten(2, 3), batch.forcing.tensor.flatten(2, 3),
            batch.outputs.tensor.flatten(2, 3), static.unsqueeze(0).expand(2, -1, -1), border, 1 - border,
            std, mean, strategy,
        )
        tgt = batch.outputs.tensor.flatten(2, 3)
        rl = ar_ref.weighted_loss(pred, tgt, torch.ones_like(tgt), w, (1 - border).squeeze(-1),
                                  (1 - border).sum().item()).mean()
        topt.zero_grad()
        rl.backward()
        if step == 0:
            flat_r = torch.cat([p.grad.flatten() for p in ref.parameters()])
            flat_g = torch.cat([p.grad.flatten().cpu() for p in lm.model.parameters()])
            assert_close(flat_g, flat_r, FP32_TOL, f"{name} rollout gradient")
        topt.step()
        assert_close(loss.cpu(), rl.detach(), FP32_TOL, f"{name} step {step} loss")
    for (n, p), (_, q) in zip(lm.model.named_parameters(), ref.named_parameters()):
        # Adam's m/sqrt(v) normalisation amplifies gradient rounding where g ~ 0 (LayerNorm biases
        # start at 0): weights are compared at 1e-3, gradients and losses above at 1e-5.
        assert_close(p.detach().cpu(), q.detach(), 1e-3, f"{name} weights after 2 steps: {n}")


def test_cuda_graph_replay_equals_eager(cuda, tmp_path):
    from py4cast_b200.datasets import synthetic_batch
    from py4cast_b200.trainer import DataParallelTrainer

    lm_a, info = _lm(tmp_path, precision="bf16-true")
    lm_b, _ = _lm(tmp_path, precision="bf16-true")
    lm_b.load_state_dict(lm_a.state_dict())
    lm_a, lm_b = lm_a.to(cuda), lm_b.to(cuda)
    batches = [synthetic_batch(info, 2, 2, 3, seed=s, device=cuda) for s in range(4)]
    eager = DataParallelTrainer(lm_a, total_steps=50)
    graph = DataParallelTrainer(lm_b, total_steps=50)
    graph.capture(batches[0], warmup=1)
    # capture ran ONE warm-up step on batches[0]: bring the eager twin to the same state
    eager.train_step(batches[0])
    assert torch.equal(eager.flat.flat, graph.flat.flat)
    for b in batches[1:]:
        le = eager.train_step(b)
        lg = graph.replay(b).clone()
        assert torch.equal(le, lg)
    assert torch.equal(eager.flat.flat, graph.flat.flat)
    assert eager.opt.host_state() == graph.opt.host_state()


def test_prefetched_host_batches_equal_direct_loads(cuda, tmp_path):
    """replay(batch, prefetch=next): the staged H2D copy on the copy stream + device-to-device move
    feeds the captured step exactly the same bytes as a direct load of the pinned host batch."""
    from py4cast_b200.datasets import synthetic_batch
    from py4cast_b200.trainer import DataParallelTrainer

    lm_a, info = _lm(tmp_path, precision="bf16-true")
    lm_b, _ = _lm(tmp_path, precision="bf16-true")
    lm_b.load_state_dict(lm_a.state_dict())
    lm_a, lm_b = lm_a.to(cuda), lm_b.to(cuda)
    host = [synthetic_batch(info, 2, 2, 3, seed=10 + s, pin=True) for s in range(4)]
    direct = DataParallelTrainer(lm_a, total_steps=50)
    staged = DataParallelTrainer(lm_b, total_steps=50)
    example = synthetic_batch(info, 2, 2, 3, seed=0, device=cuda)
    direct.capture(example, warmup=1)
    staged.capture(example, warmup=1)
    staged.prefetch(host[0])
    for i, hb in enumerate(host):
        ld = direct.replay(hb).clone()
        ls = staged.replay(hb, prefetch=host[i + 1] if i + 1 < len(host) else None).clone()
        assert torch.equal(ld, ls)
    # a batch that was not announced is loaded directly
    assert torch.equal(direct.replay(host[1]).clone(), staged.replay(host[1]).clone())
    assert torch.equal(direct.flat.flat, staged.flat.flat)


def test_launch_counter_counts_our_kernels(cuda, tmp_path):
    from py4cast_b200 import _lib
    from py4cast_b200.datasets import synthetic_batch
    from py4cast_b200.trainer import DataParallelTrainer

    lm, info = _lm(tmp_path)
    tr = DataParallelTrainer(lm.to(cuda))
    b = synthetic_batch(info, 2, 2, 3, device=cuda)
    tr.train_step(b)
    _lib.reset_launch_count()
    tr.train_step(b)
    assert _lib.launch_count() > 100


def test_stale_deferred_wgrads_do_not_leak_into_the_next_step(cuda, tmp_path):
    """A backward outside the trainer (sinks installed, never flushed) leaves weight-gradient
    partials behind; the next trainer step must overwrite them, not accumulate onto them."""
    from py4cast_b200.datasets import synthetic_batch
    from py4cast_b200.trainer import DataParallelTrainer

    lm, info = _lm(tmp_path, precision="bf16-true")
    lm = lm.to(cuda)
    tr = DataParallelTrainer(lm, total_steps=10)
    batch = synthetic_batch(info, 2, 2, 3, seed=5, device=cuda)
    tr._forward_backward(_fresh_batch(batch))
    clean = tr.flat.grad.clone()
    # stray backward: partials written, flush never called
    lm.training_step(_fresh_batch(batch), 1).backward()
    lm.training_step_losses.clear()
    tr._forward_backward(_fresh_batch(batch))
    assert torch.equal(tr.flat.grad, clean)


def test_inference_after_training_with_the_trainer(cuda, tmp_path):
    """The trainer's first step records the feature / dim names (reference lightning.py:541-545),
    so `forward()` (phase="inference") works afterwards."""
    from py4cast_b200.datasets import ItemBatch, synthetic_batch
    from py4cast_b200.trainer import DataParallelTrainer

    lm, info = _lm(tmp_path, precision="bf16-true")
    lm = lm.to(cuda)
    tr = DataParallelTrainer(lm, total_steps=10)
    batch = synthetic_batch(info, 2, 2, 3, seed=5, device=cuda)
    tr.train_step(batch)
    assert lm.output_feature_names == batch.outputs.feature_names
    fb = _fresh_batch(batch)
    with torch.no_grad():
        pred = lm(ItemBatch(inputs=fb.inputs, forcing=fb.forcing, outputs=None), 1)
    assert pred.tensor.shape[:2] == batch.outputs.tensor.shape[:2]
    assert torch.isfinite(pred.tensor).all()


def _fresh_batch(batch):
    from py4cast_b200.trainer import _fresh

    return _fresh(batch)


def test_host_batch_input_pipeline_fills_the_captured_buffers(cuda, tmp_path):
    """The slimmed host batch (NWP forcing channels in bf16, external forcings as 28 bytes of scalars per
    slot) lands in the captured fp32 input buffers exactly as the full fp32 batch would: shipped channels
    bit-equal to their bf16 rounding (what the bf16 `x` holds anyway), the five external channels equal to
    `fill_external_forcings`, inputs / outputs untouched - directly and through the prefetch stream."""
    import datetime as dt

    from py4cast_b200.datasets import synthetic_batch, synthetic_dataset_info
    from py4cast_b200.forcingutils import fill_external_forcings
    from py4cast_b200.lightning import AutoRegressiveLightning
    from py4cast_b200.trainer import DataParallelTrainer, HostBatch

    H, W, T, B = 64, 64, 3, 2
    info = synthetic_dataset_info(H, W, num_features=5, num_forcing=8, border_size=4)
    torch.manual_seed(0)
    lm = AutoRegressiveLightning(
        settings_init_args={"tmp_dir": str(tmp_path), "processor_layers": 1}, dataset_info=info, num_input_steps=2,
        num_pred_steps_train=T, batch_size=B, model_name="GraphLAM", training_strategy="diff_ar",
        precision="bf16-true").to(cuda)
    tr = DataParallelTrainer(lm, total_steps=100)
    lat = torch.linspace(51, 41, H, device=cuda).view(H, 1).expand(H, W).contiguous()
    lon = torch.linspace(-6, 10, W, device=cuda).view(1, W).expand(H, W).contiguous()
    tr.configure_external_forcings(lat, lon)
    dates = [dt.datetime(2023, 12, 31, 21), dt.datetime(2024, 6, 1, 6)]
    leads = [dt.timedelta(hours=h + 1) for h in range(T)]
    batches = [synthetic_batch(info, B, 2, T, seed=s, device=cuda) for s in (1, 2)]
    expected = []
    for b in batches:
        f = b.forcing.tensor.clone()
        f[..., :3] = f[..., :3].bfloat16().float()
        fill_external_forcings(f, 3, lat, lon, dates, leads)
        expected.append(f)
    tr.capture(batches[0], warmup=1)
    hbs = [tr.pack_host_batch(b, dates, leads) for b in batches]
    assert isinstance(hbs[0], HostBatch) and hbs[0].forcing.dtype == torch.bfloat16 and hbs[0].forcing.shape[-1] == 3
    assert hbs[0].nbytes() < sum(getattr(batches[0], n).tensor.numel() * 4 for n in ("inputs", "forcing", "outputs"))
    tr.load_inputs(hbs[1])
    torch.cuda.synchronize()
    assert torch.equal(tr._static_batch.forcing.tensor, expected[1])
    assert torch.equal(tr._static_batch.inputs.tensor, batches[1].inputs.tensor)
    assert torch.equal(tr._static_batch.outputs.tensor, batches[1].outputs.tensor)
    # through the copy stream: step on batch 0 while batch 1 is prefetched, then step on batch 1
    l0 = tr.replay(hbs[0], prefetch=hbs[1]).clone()
    torch.cuda.synchronize()
    assert torch.equal(tr._static_batch.forcing.tensor, expected[0])
    l1 = tr.replay(hbs[1]).clone()
    torch.cuda.synchronize()
    assert torch.equal(tr._static_batch.forcing.tensor, expected[1])
    assert torch.isfinite(l0) and torch.isfinite(l1)
