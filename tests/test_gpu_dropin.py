"""Drop-in proof: the registered CUDA model and the CUDA losses plugged into the REFERENCE's own
AR loop.

Two forms of the same experiment (reference: py4cast/lightning.py:265-306 construction,
:495-676 `_common_step`, :806-831 `training_step`; py4cast/models.py:66-89):

  * `test_reference_loop_restated_...`: the loop is oracle/ar_ref.ar_rollout — the restatement that
    tests/test_oracle_pinning.py pins line by line against the verbatim reference file — executed with
    stock torch ops on the GPU, with `py4cast_b200.nlam.HiLAM` as `model` and
    `py4cast_b200.losses.WeightedLoss` as the loss.  This runs on the GPU box.
  * `test_verbatim_reference_lightning_...`: the reference's UNMODIFIED lightning.py, loaded by
    oracle/refshim.py, with `self.model = py4cast_b200.nlam.HiLAM`, the product NamedTensor and the
    product CombinedLoss.  It needs both a GPU and /root/reference and is skipped where either is
    missing (the build container has no GPU, the GPU box has no reference tree).

Both compare with `py4cast_b200.lightning.AutoRegressiveLightning` (the fused path) on the same
batch: prediction, loss, parameter gradients, fp32 at 1e-5.
"""
import pytest
import torch

from helpers import FP32_TOL, assert_close

pytestmark = pytest.mark.gpu

H, W = 64, 80
LOSS = [{"class": "WeightedLoss", "weight": 1.0, "params": {"loss": "MSELoss", "reduction": "none"}}]


def _product(tmp_path, cuda, strategy, T):
    from py4cast_b200.datasets import synthetic_batch, synthetic_dataset_info
    from py4cast_b200.lightning import AutoRegressiveLightning

    info = synthetic_dataset_info(H, W, num_features=5, num_forcing=3, border_size=4)
    torch.manual_seed(3)
    lm = AutoRegressiveLightning(
        settings_init_args={"tmp_dir": str(tmp_path), "processor_layers": 2}, dataset_info=info,
        num_input_steps=2, num_pred_steps_train=T, batch_size=2, model_name="HiLAM", losses=LOSS,
        training_strategy=strategy, precision="32-true",
    ).to(cuda)
    batch = synthetic_batch(info, 2, 2, T, seed=21, device=cuda)
    return lm, info, batch


def _grads(model):
    g = torch.cat([p.grad.flatten().float() for p in model.parameters()]).clone()
    model.zero_grad(set_to_none=True)
    return g


@pytest.mark.parametrize("strategy", ["scaled_ar", "diff_ar"])
def test_reference_loop_restated_with_cuda_model_and_loss(cuda, tmp_path, strategy):
    from oracle import ar_ref
    from py4cast_b200.namedtensor import NamedTensor
    from py4cast_b200.trainer import _fresh

    T = 3
    lm, info, batch = _product(tmp_path, cuda, strategy, T)
    # fused path
    with torch.no_grad():
        pred_f, _ = lm.common_step(_fresh(batch), 0, "train")
    loss_f = lm.training_step(_fresh(batch), 0)
    loss_f.backward()
    g_f = _grads(lm.model)

    # the reference's loop (pinned restatement), stock torch ops around the registered model
    fn = info.shortnames["input_output"]
    std = info.diff_stats.to_list("std", fn).to(cuda)
    mean = info.diff_stats.to_list("mean", fn).to(cuda)
    border = info.statics.border_mask.flatten(0, 1).to(cuda)
    static = info.statics.grid_statics.tensor.flatten(0, 1).unsqueeze(0).expand(2, -1, -1).to(cuda)
    b = _fresh(batch)
    tgt = b.outputs.tensor.flatten(2, 3)
    pred_r = ar_ref.ar_rollout(
        lm.model, b.inputs.tensor.flatten(2, 3), b.forcing.tensor.flatten(2, 3), tgt, static, border,
        1 - border, std, mean, strategy,
    )
    names = ["batch", "timestep", "ngrid", "features"]
    loss_r = lm.loss(NamedTensor(pred_r, names, list(fn)), NamedTensor(tgt, names, list(fn)),
                     torch.ones_like(tgt)).mean()
    loss_r.backward()
    g_r = _grads(lm.model)

    assert_close(pred_f.tensor.reshape(pred_r.shape), pred_r.detach(), FP32_TOL, f"drop-in {strategy} prediction")
    assert_close(loss_f.detach(), loss_r.detach(), FP32_TOL, f"drop-in {strategy} loss")
    assert_close(g_f, g_r, 10 * FP32_TOL, f"drop-in {strategy} parameter gradients")


def test_verbatim_reference_lightning_with_cuda_model_and_loss(cuda, tmp_path):
    from oracle import refshim

    if not refshim.reference_available():
        pytest.skip("needs /root/reference next to a GPU (absent on the GPU box)")
    from py4cast_b200.namedtensor import NamedTensor
    from py4cast_b200.trainer import _fresh

    T = 3
    lm, info, batch = _product(tmp_path, cuda, "scaled_ar", T)
    with torch.no_grad():
        pred_f, _ = lm.common_step(_fresh(batch), 0, "train")
    loss_f = lm.training_step(_fresh(batch), 0)

    fn = info.shortnames["input_output"]
    diff = refshim.DictStats({n: {"std": info.diff_stats[n]["std"], "mean": info.diff_stats[n]["mean"]} for n in fn})
    border = info.statics.border_mask.flatten(0, 1).to(cuda)
    static = info.statics.grid_statics.tensor.flatten(0, 1).unsqueeze(0).expand(2, -1, -1).contiguous().to(cuda)
    ref_lm = refshim.make_reference_module(
        lm.model, training_strategy="scaled_ar", border_mask=border, grid_static_features=static,
        diff_stats=diff, precision="32-true", named_tensor_cls=NamedTensor,
    )
    ref_lm.loss = lm.loss  # product CombinedLoss, prepared by the product module
    pred_r, target_r = ref_lm.common_step(_fresh(batch), 0, "train")
    assert_close(pred_f.tensor, pred_r.tensor.detach(), FP32_TOL, "verbatim lightning.py + CUDA HiLAM: prediction")
    loss_r = ref_lm.loss(pred_r, target_r, torch.ones_like(target_r.tensor)).mean()
    assert_close(loss_f.detach(), loss_r.detach(), FP32_TOL, "verbatim lightning.py + CUDA losses: loss")
