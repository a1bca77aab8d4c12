"""Model registry / plugin discovery: mirrors reference tests/test_models.py:145-165 and the
error behaviour of py4cast/models.py:42-45, 55-62."""
import importlib
import sys
import textwrap

import pytest

from py4cast_b200 import models
from py4cast_b200.base import ModelABC, ModelType
from py4cast_b200.nlam import GraphLamSettings


def test_registry_has_the_graph_family():
    assert {"HiLAM", "GraphLAM", "HiLAMParallel"} <= set(models.registry)
    assert "PanguWeather" not in models.registry and "ArchesWeather" not in models.registry
    for name in ("HiLAM", "GraphLAM", "HiLAMParallel"):
        kls = models.registry[name]
        assert issubclass(kls, ModelABC) and kls.model_type == ModelType.GRAPH
        assert kls.features_last is True and kls.onnx_supported is False
        assert kls.num_spatial_dims == 1 and kls.supported_num_spatial_dims == (1,)
        assert kls.settings_kls is GraphLamSettings and hasattr(kls, "rank_zero_setup")


def test_settings_defaults_and_errors():
    kls, s = models.get_model_kls_and_settings("HiLAM", {})
    assert (s.hidden_dims, s.hidden_layers, s.processor_layers, s.mesh_aggr) == (64, 1, 4, "sum")
    assert s.use_checkpointing is False and s.offload_to_cpu is False
    _, s2 = models.get_model_kls_and_settings("GraphLAM", {"processor_layers": 2, "mesh_aggr": "mean"})
    assert s2.processor_layers == 2 and s2.to_dict()["mesh_aggr"] == "mean"
    with pytest.raises(TypeError):
        models.get_model_kls_and_settings("HiLAM", {"not_a_field": 1})
    with pytest.raises(KeyError, match="not found in registry"):
        models.get_model_kls_and_settings("GraphLam", {})  # the reference's own yaml typo


PLUGIN = """
from dataclasses import dataclass
import torch
from torch import nn
from py4cast_b200.base import ModelABC, ModelType

@dataclass(slots=True)
class IdentitySettings:
    name: str = "Identity"

class {cls}(ModelABC, nn.Module):
    settings_kls = IdentitySettings
    onnx_supported = False
    features_last = True
    supported_num_spatial_dims = (2,)
    num_spatial_dims = 2
    model_type = ModelType.CONVOLUTIONAL
    register = {register}

    def __init__(self, in_channels, out_channels, input_shape, settings, *args, **kwargs):
        super().__init__()
        self.in_channels, self.out_channels, self.input_shape = in_channels, out_channels, input_shape
        self._settings = settings
        self.scaler = nn.Parameter(torch.rand(1))
        self.check_required_attributes()

    @property
    def settings(self):
        return self._settings

    def forward(self, x):
        return x[..., : self.out_channels] * self.scaler
"""


def _write_plugin(tmp_path, modname, cls, register=True):
    (tmp_path / f"{modname}.py").write_text(textwrap.dedent(PLUGIN.format(cls=cls, register=register)))


def test_plugin_discovery(tmp_path, monkeypatch):
    _write_plugin(tmp_path, "py4cast_plugin_tmp_identity", "Identity")
    _write_plugin(tmp_path, "py4cast_plugin_tmp_hidden", "Hidden", register=False)
    monkeypatch.syspath_prepend(str(tmp_path))
    importlib.invalidate_caches()
    reg = dict(models.registry)
    models.discover_plugins(reg)
    assert "Identity" in reg and "Hidden" not in reg
    model = reg["Identity"](3, 2, (8, 8), reg["Identity"].settings_kls())
    assert model.features_second is False
    for m in ("py4cast_plugin_tmp_identity", "py4cast_plugin_tmp_hidden"):
        sys.modules.pop(m, None)


def test_plugin_duplicate_name_raises(tmp_path, monkeypatch):
    _write_plugin(tmp_path, "py4cast_plugin_tmp_dup", "HiLAM")
    monkeypatch.syspath_prepend(str(tmp_path))
    importlib.invalidate_caches()
    with pytest.raises(ValueError, match="already exists in the registry"):
        models.discover_plugins(dict(models.registry))
    sys.modules.pop("py4cast_plugin_tmp_dup", None)


def test_graph_model_needs_rank_zero_setup_first(tmp_path):
    kls, s = models.get_model_kls_and_settings("GraphLAM", {"tmp_dir": str(tmp_path)})
    with pytest.raises(FileNotFoundError, match="rank_zero_setup"):
        kls(2, 1, (64, 64), s)


def test_unsupported_settings_fail_loudly(tmp_path):
    from helpers import pixel_meshgrid

    for bad in ({"hidden_layers": 2}, {"hidden_dims": 32}):
        kls, s = models.get_model_kls_and_settings("GraphLAM", {"tmp_dir": str(tmp_path), **bad})
        kls.rank_zero_setup(s, pixel_meshgrid(64, 64))
        with pytest.raises(NotImplementedError):
            kls(in_channels=2, out_channels=1, settings=s, input_shape=(64, 64))


def test_memory_knobs_of_the_reference_are_accepted(tmp_path):
    """`use_checkpointing` / `offload_to_cpu` (graphlam.yaml:23-24) move activations, never change a result: a
    configuration that sets them builds (with a warning) instead of failing."""
    from helpers import pixel_meshgrid

    for knob in ({"use_checkpointing": True}, {"offload_to_cpu": True}):
        kls, s = models.get_model_kls_and_settings("GraphLAM", {"tmp_dir": str(tmp_path), **knob})
        kls.rank_zero_setup(s, pixel_meshgrid(64, 64))
        with pytest.warns(UserWarning, match="no effect"):
            m = kls(in_channels=2, out_channels=1, settings=s, input_shape=(64, 64))
        assert m.settings is s
