"""
Model registry and plugin discovery — same names, call signatures and error
behaviour as reference py4cast/models.py:16-89:

  * `registry`: class name -> class.  Starts from mfai's registry when mfai is importable
    (minus PanguWeather / ArchesWeather, models.py:19-20), with the graph family
    (GraphLAM / HiLAM / HiLAMParallel) provided by THIS package's B200 kernels;
  * every `ModelABC` subclass with `register = True` found in an importable top-level module
    named `py4cast_plugin_*` is added (models.py:23-46); a duplicate name raises ValueError;
  * `get_model_kls_and_settings(name, settings_dict)` -> (class, settings instance),
    KeyError with the registry's message for unknown names (models.py:50-63);
  * `build_model_from_settings(name, n_in, n_out, settings_dict, input_shape)` calls
    `kls(n_in, n_out, input_shape, settings)` positionally (models.py:66-89).
"""

from __future__ import annotations

import importlib
import pkgutil
from typing import Any, Tuple

from .base import ModelABC
from .nlam import GRAPH_MODELS
from .standins import HalfUNet

registry: dict = {}
_mfai_model_abc = None
try:  # optional: the reference's own third-party model zoo
    from mfai.pytorch.models import registry as _mfai_registry  # type: ignore
    from mfai.pytorch.models.base import ModelABC as _mfai_model_abc  # type: ignore

    registry.update(_mfai_registry)
    registry.pop("PanguWeather", None)
    registry.pop("ArchesWeather", None)
except ImportError:
    registry["HalfUNet"] = HalfUNet  # stock-PyTorch stand-in, see standins.py
# the graph family always runs on the B200 kernels
registry.update(GRAPH_MODELS)

PLUGIN_PREFIX = "py4cast_plugin_"


def _is_model_class(kls) -> bool:
    bases = tuple(b for b in (ModelABC, _mfai_model_abc) if b is not None)
    return isinstance(kls, type) and issubclass(kls, bases) and kls not in bases and bool(kls.register)


def discover_plugins(target: dict) -> dict:
    discovered = {}
    for _, name, _ in pkgutil.iter_modules():
        if name.startswith(PLUGIN_PREFIX):
            discovered[name] = importlib.import_module(name)
    for module_name, module in discovered.items():
        for _, kls in module.__dict__.items():
            if _is_model_class(kls):
                if kls.__name__ in target:
                    raise ValueError(
                        f"Model {kls.__name__} from plugin {module_name} already exists in the registry."
                    )
                target[kls.__name__] = kls
    return discovered


discovered_modules = discover_plugins(registry)
all_nn_architectures = list(registry)


def get_model_kls_and_settings(model_name: str, settings: dict):
    """Returns the classes for a model and its settings instance."""
    try:
        model_kls = registry[model_name]
    except KeyError as e:
        raise KeyError(
            f"Model {model_name} not found in registry of {__file__}. Did you add it ? Names are {registry.keys()}"
        ) from e
    settings_kls = model_kls.settings_kls
    model_settings = settings_kls(**settings)
    return model_kls, model_settings


def build_model_from_settings(
    network_name: str,
    num_input_features: int,
    num_output_features: int,
    settings: dict,
    input_shape: tuple,
    *args,
    **kwargs,
) -> Tuple[ModelABC, Any]:
    """Instanciates a model based on its name and an optional settings dict."""
    model_kls, model_settings = get_model_kls_and_settings(network_name, settings)
    return (
        model_kls(num_input_features, num_output_features, input_shape, model_settings, *args, **kwargs),
        model_settings,
    )
