"""Host-side contract checks that need no GPU: bench.py's defaults and workload description, the
best-effort NUMA binding, and the loud failure of the evaluation steps on CPU tensors."""
import importlib.util
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_bench_defaults_and_workload_description(monkeypatch):
    bench = _bench()
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    a = bench.parse_args()
    assert (a.gpus, a.impl, a.model, a.pred_steps, a.precision, a.strategy) == (1, "b200", "HiLAM", 4, "bf16-true", "diff_ar")
    assert a.batch_per_gpu == 16 and a.warmup >= 3
    cfg = bench.workload_config(a, 8)
    assert cfg["global_batch"] == 128 and cfg["parallelism"] == "dp8" and cfg["grid"] == [512, 640]
    assert "HiLAM" in cfg["workload"] and "model" not in cfg  # config names the workload, no model keys
    monkeypatch.setattr(sys, "argv", ["bench.py", "--pred-steps", "12"])
    assert bench.parse_args().batch_per_gpu == 5  # 64 // T for long rollouts
    monkeypatch.setattr(sys, "argv", ["bench.py", "--batch-per-gpu", "2"])
    assert bench.parse_args().batch_per_gpu == 2
    assert bench.METRIC == "hilam_ar_rollout_train_samples_per_s" and bench.UNIT == "samples/s"


def test_numa_binding_is_best_effort_without_a_gpu():
    from py4cast_b200.trainer import bind_to_gpu_numa_node

    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    before = os.sched_getaffinity(0)
    assert bind_to_gpu_numa_node(0) is False
    assert os.sched_getaffinity(0) == before


def test_evaluation_steps_refuse_cpu_batches(tmp_path):
    """No CPU fallback: validation_step / test_step on a host batch raise, like training_step."""
    from py4cast_b200 import _lib
    from py4cast_b200.datasets import synthetic_batch, synthetic_dataset_info
    from py4cast_b200.lightning import AutoRegressiveLightning

    info = synthetic_dataset_info(64, 64, num_features=3, num_forcing=2, border_size=2)
    lm = AutoRegressiveLightning({"tmp_dir": str(tmp_path), "processor_layers": 1}, info, num_input_steps=2,
                                 num_pred_steps_train=1, num_pred_steps_val_test=2, batch_size=1, model_name="HiLAM")
    for step in (lm.validation_step, lm.test_step):
        with pytest.raises(_lib.P4CError):
            step(synthetic_batch(info, 1, 2, 2, seed=1), 0)
