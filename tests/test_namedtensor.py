"""NamedTensor mirror: API pinned by the reference call sites (tests/test_datasets.py:15-111,
doc/features.md:134-164) + the verbatim reference hot path running on top of it."""
import pytest
import torch

from py4cast_b200.datasets import Item, ItemBatch, collate_fn
from py4cast_b200.namedtensor import NamedTensor


def _nt(*shape, names, feats):
    return NamedTensor(torch.rand(*shape), names=names, feature_names=feats)


def test_doc_examples():
    nt = _nt(4, 256, 256, 3, names=["batch", "lat", "lon", "features"], feats=["u", "v", "t2m"])
    assert nt.dim_size("lat") == 256
    nt2 = _nt(4, 256, 256, 1, names=["batch", "lat", "lon", "features"], feats=["q"])
    nt3 = nt | nt2
    assert nt3.feature_names == ["u", "v", "t2m", "q"] and nt3.tensor.shape[-1] == 4
    assert nt3["u"].shape == (4, 256, 256, 1)
    nt4 = NamedTensor.new_like(torch.rand(4, 256, 256, 4), nt3)
    assert nt4.names == nt3.names
    assert nt3.spatial_dim_idx == [1, 2] and nt3.num_spatial_dims == 2
    nt3.flatten_("ngrid", 1, 2)
    assert nt3.names == ["batch", "ngrid", "features"] and nt3.tensor.shape == (4, 256 * 256, 4)
    assert nt3.spatial_dim_idx == [1]
    nt3.unflatten_(1, (256, 256), ["lat", "lon"])
    assert nt3.names == ["batch", "lat", "lon", "features"]
    assert "Feature name" in str(nt)


def test_selection_and_errors():
    nt = _nt(2, 3, 5, 4, names=["batch", "timestep", "ngrid", "features"], feats=list("abcd"))
    assert nt.select_tensor_dim("timestep", -1).shape == (2, 5, 4)
    sub = nt.select_dim("timestep", 0)
    assert sub.names == ["batch", "ngrid", "features"]
    assert nt.index_select_tensor_dim("timestep", range(1, 3)).shape == (2, 2, 5, 4)
    assert nt.dim_index("ngrid") == 2 and nt.dim_size("nope") is None
    c = nt.clone()
    c.tensor.zero_()
    assert nt.tensor.abs().sum() > 0
    with pytest.raises(ValueError):
        NamedTensor(torch.rand(2, 3), ["a", "features"], ["x"])
    with pytest.raises(ValueError):
        NamedTensor(torch.rand(2, 3), ["a"], ["x", "y", "z"])
    with pytest.raises(ValueError):
        nt | _nt(2, 3, 5, 1, names=["batch", "timestep", "lat", "features"], feats=["e"])
    with pytest.raises(ValueError):
        nt | _nt(2, 3, 5, 1, names=["batch", "timestep", "ngrid", "features"], feats=["a"])
    e = NamedTensor.expand_to_batch_like(torch.rand(7, 2, 3, 5, 4), nt)
    assert e.names[0] == "batch" and len(e.names) == 5


def test_item_validation_and_collate():
    """reference tests/test_datasets.py:15-111"""
    names = ["timestep", "lat", "lon", "features"]
    mk = lambda f: NamedTensor(torch.rand(2, 8, 8, len(f)).double(), names, f)
    item = Item(inputs=mk(["a", "b"]), forcing=mk(["f"]), outputs=mk(["a", "b"]))
    batch = collate_fn([item, item, item])
    assert isinstance(batch, ItemBatch)
    assert batch.inputs.names == ["batch"] + names and batch.inputs.tensor.dtype == torch.float32
    assert batch.batch_size == 3 and batch.num_input_steps == 2 and batch.num_pred_steps == 2
    with pytest.raises(ValueError):
        Item(inputs=mk(["a", "b"]), forcing=mk(["f"]), outputs=mk(["a", "c"]))
    with pytest.raises(ValueError):
        Item(inputs=NamedTensor(torch.rand(2, 8, 8, 2), ["t", "lat", "lon", "features"], ["a", "b"]),
             forcing=mk(["f"]), outputs=mk(["a", "b"]))


def test_verbatim_reference_runs_on_product_namedtensor():
    """Drop-in check: the reference's own lightning.py / losses.py execute unmodified on top of
    py4cast_b200.NamedTensor and give the golden answer."""
    from oracle import make_golden, refshim

    if not refshim.reference_available():
        pytest.skip("needs /root/reference")
    from helpers import assert_close, load_golden
    from oracle import mfai_shim

    case = load_golden("ar_rollout.pt")["graph_scaled_ar"]
    p = make_golden.make_problem(1234, B=2, T_in=2, T=4, H=16, W=20, F=3, Ff=2, border=2)
    _, ref_lightning = refshim.load_reference(NamedTensor)
    model = make_golden.LinearTanh(2 * 3 + 4 + 2, 3, True, False, 7)
    info = make_golden.dataset_info_for(p)
    lm = refshim.make_reference_module(
        model, training_strategy="scaled_ar", border_mask=p.border_mask.flatten(0, 1),
        grid_static_features=mfai_shim.expand_to_batch(p.statics.flatten(0, 1), 2),
        diff_stats=info.diff_stats, named_tensor_cls=NamedTensor,
    )
    batch = ItemBatch(
        inputs=NamedTensor(p.inputs.clone(), make_golden.NAMES5, p.feature_names),
        forcing=NamedTensor(p.forcing.clone(), make_golden.NAMES5, p.forcing_names),
        outputs=NamedTensor(p.outputs.clone(), make_golden.NAMES5, p.feature_names),
    )
    pred, _ = lm.common_step(batch, 0, "train")
    assert pred.names == ["batch", "timestep", "ngrid", "features"]
    assert_close(pred.tensor.reshape(2, 4, 16, 20, 3), case["expected"]["pred"], 1e-6)
