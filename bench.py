#!/usr/bin/env python
"""
HiLAM AR-rollout training throughput on B200 (BASELINE.json metric), one process per GPU.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the CPU arm (oracle port of the reference path)

A "step" = one optimiser step on one synthetic Titan-shaped batch per GPU: T-step AR rollout
(fwd) + WeightedLoss(MSE) + backward + gradient all-reduce + AdamW (accumulate_grad_batches = 1).
Workload = BASELINE configs[2] (HiLAM, 3 mesh levels, 512x640 grid, F=21, F_f=21, F_s=4,
T_in=2, diff_ar, T=4, bf16); `--pred-steps 12` gives configs[3].  Batch 16 per GPU at T <= 4 (64 / T
beyond): the reference trains this grid with batch_size 2 on 32-GB parts (config/CLI/dataset/
titan.yaml:7, with activation checkpointing); 180 GB of HBM3e hold 16 samples' activations (~3 GB
each at T=4) with room, and the larger batch amortises the ~14 ms of launch-bound mesh-level kernels
per step (profiles/r01f_batch_sweep.txt, r01t_batch_sweep.txt: 64.5 / 83.5 / 96.8 samples/s at batch
2 / 4 / 8 early in the round, 108 / 113 / 119 at batch 8 / 12 / 16 at its end).
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the ONE JSON line (NCCL prints its version there)

import torch  # noqa: E402

METRIC = "hilam_ar_rollout_train_samples_per_s"
UNIT = "samples/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--pred-steps", type=int, default=4)
    # reference config/CLI/dataset/titan.yaml:7: batch_size 2 per (32 GB) GPU.  0 = auto: 16 for
    # rollouts of up to 4 steps (~3 GB of saved activations per sample), 64 // T for longer ones
    ap.add_argument("--batch-per-gpu", type=int, default=0)
    ap.add_argument("--precision", default="bf16-true", choices=["bf16-true", "32-true"])
    ap.add_argument("--model", default="HiLAM", choices=["HiLAM", "GraphLAM", "HiLAMParallel"])
    ap.add_argument("--strategy", default="diff_ar", choices=["diff_ar", "scaled_ar"])
    ap.add_argument("--grid", default="512x640")
    ap.add_argument("--processor-layers", type=int, default=4,
                    help="mesh processor depth (reference hilam.yaml: 4); other values are for experiments only")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=120.0)
    ap.add_argument("--profile-step", action="store_true",
                    help="run ONE eager step between cudaProfilerStart/Stop after warm-up and exit "
                         "(for `ncu --profile-from-start off`; prints no bench line)")
    args = ap.parse_args()
    if args.batch_per_gpu <= 0:
        args.batch_per_gpu = 16 if args.pred_steps <= 4 else max(1, 64 // args.pred_steps)
    return args


def workload_config(args, world):
    H, W = (int(v) for v in args.grid.split("x"))
    return {
        "workload": f"{args.model} 3-level mesh, {H}x{W} grid, F=21 F_forcing=21 F_static=4, T_in=2, "
                    f"{args.strategy} T={args.pred_steps}, WeightedLoss(MSE), AdamW, batch {args.batch_per_gpu}/GPU",
        "model_name": args.model, "grid": [H, W], "pred_steps": args.pred_steps,
        "batch_per_gpu": args.batch_per_gpu, "global_batch": args.batch_per_gpu * world,
        "strategy": args.strategy, "precision": args.precision, "parallelism": f"dp{world}",
        "l2": "per-step working set (several GB of activations) >> 126 MB L2: no flush needed",
    }


# ------------------------------------------------------------------ clocks sampling
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------ CPU arm
def cpu_reference_arm(args, cfg, steps, warmup, budget_s):
    """Oracle port of the reference path (pure-torch HiLAM restatement + AR/loss restatement +
    torch.optim.AdamW, fp32 — the reference's CPU path never enters autocast) on all host cores.
    Bounded sample: each timed step is ONE AR step of the rollout (T=1) at the full grid and
    batch 1; samples/s = 1 / (T * seconds per AR step)."""
    from oracle import ar_ref, graph_ref
    from py4cast_b200.datasets import synthetic_batch, synthetic_dataset_info
    from py4cast_b200.graph_build import build_mesh_graph

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    H, W = cfg["grid"]
    info = synthetic_dataset_info(H, W)
    hier = args.model != "GraphLAM"
    graph = build_mesh_graph(info.statics.meshgrid, hier, 3 if hier else None)
    torch.manual_seed(0)
    f_in = 2 * info.weather_dim + 4 + info.forcing_dim
    model = getattr(graph_ref, args.model)(f_in, info.weather_dim, graph)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, betas=(0.9, 0.95))
    fn = info.shortnames["input_output"]
    std, mean = info.diff_stats.to_list("std", fn), info.diff_stats.to_list("mean", fn)
    w = torch.tensor([info.state_weights[n] for n in fn]) / std**2
    border = info.statics.border_mask.flatten(0, 1)
    static = info.statics.grid_statics.tensor.flatten(0, 1).unsqueeze(0)
    batch = synthetic_batch(info, 1, 2, 1, seed=0)
    tgt = batch.outputs.tensor.flatten(2, 3)
    ones = torch.ones_like(tgt)

    def one_step():
        pred = ar_ref.ar_rollout(
            model, batch.inputs.tensor.flatten(2, 3), batch.forcing.tensor.flatten(2, 3), tgt, static, border,
            1 - border, std, mean, args.strategy,
        )
        loss = ar_ref.weighted_loss(pred, tgt, ones, w, (1 - border).squeeze(-1), (1 - border).sum().item()).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()

    t_start = time.perf_counter()
    for _ in range(max(1, min(warmup, 1))):
        one_step()
    times = []
    for _ in range(max(1, steps)):
        t0 = time.perf_counter()
        one_step()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s and len(times) >= 1:
            break
    sec_per_ar_step = sum(times) / len(times)
    T = cfg["pred_steps"]
    value = 1.0 / (T * sec_per_ar_step)
    return {
        "value": value, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": f"{len(times)} timed optimiser steps of a 1-AR-step rollout (fwd+loss+bwd+AdamW, fp32, "
                  f"batch 1, full {H}x{W} grid, {sec_per_ar_step:.2f} s each) after 1 warm-up; samples/s = 1/(T*t), T={T}",
        "timed_steps": len(times), "sec_per_ar_step": sec_per_ar_step,
    }


# --------------------------------------------------------------- kernel roofline
def kernel_roofline(lm, info, cfg, dev):
    """Times the message-passing kernels of the m2g InteractionNet (the largest edge set:
    1.31 M edges, 327 680 receivers) in isolation with CUDA events on the launching stream and
    reports achieved algorithmic GB/s against the measured HBM peak."""
    import ctypes as C

    from py4cast_b200 import _lib, ops
    from py4cast_b200._lib import Mlp, check, ptr, stream

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    net = lm.model.m2g_gnn
    gi = net.graph_index()
    dt = torch.bfloat16 if cfg["precision"] == "bf16-true" else torch.float32
    s = 2 if dt == torch.bfloat16 else 4
    B, E, Ns, Nr, d = cfg["batch_per_gpu"], gi.num_edges, gi.num_send, gi.num_rec, 64
    g = torch.Generator(device="cuda").manual_seed(0)
    send = torch.randn(B, Ns, d, device=dev, generator=g).to(dt)
    rec = torch.randn(B, Nr, d, device=dev, generator=g).to(dt)
    edge = torch.randn(E, d, device=dev, generator=g).to(dt)
    h1 = torch.empty(B, E, d, device=dev, dtype=dt)
    st = torch.empty(B, E, 2, device=dev, dtype=torch.float32)
    aggr = torch.empty(B, Nr, d, device=dev, dtype=dt)
    msg = torch.empty(B, E, d, device=dev, dtype=dt)
    lib = _lib.load()
    lm.model.pack_weights()
    desc = Mlp()
    ops._fill_mlp(desc, E, B, E, [ops._SrcSpec(edge), ops._SrcSpec(send, index=gi.send_idx),
                                  ops._SrcSpec(rec, index=gi.rec_idx)], net.edge_mlp.weights(), d, dt,
                  net.edge_mlp.meta)
    fused = dt == torch.bfloat16 and gi.uniform_deg > 0

    def edge_fwd():
        if fused:  # aggregation in the epilogue: the messages never reach HBM
            check(lib.p4c_mlp_forward_aggregate(C.byref(desc), None, None, ptr(h1), ptr(st), ptr(aggr),
                                                gi.uniform_deg, 1.0, stream()), "edge fwd")
        else:
            check(lib.p4c_mlp_forward(C.byref(desc), None, ptr(msg), None, ptr(h1), ptr(st), stream()), "edge fwd")

    def timeit(fn, iters=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e-3

    out = []
    t = timeit(edge_fwd)
    # s*d*(N_send + N_rec + E) read (sender rows, receiver rows, edge reps) + saved
    # activations written (h1: s*d*E, LN stats: 8*E) + index bytes (8*E) + the output: s*d*N_rec
    # aggregated rows (fused) or s*d*E messages (unfused)
    bytes_edge = B * (s * d * (Ns + Nr + E) + s * d * E + 8 * E + (s * d * Nr if fused else s * d * E)) + 8 * E
    name = ("mlp2_ws_fwd_kernel" if fused else "mlp2_tc_fwd_kernel") + " (m2g edge MLP: gather+MLP+LN" + (
        "+aggregate)" if fused else ")")
    out.append({"kernel": name, "bytes": bytes_edge, "seconds": t})
    if not fused:
        t = timeit(lambda: ops.segment_sum(msg, gi.by_rec, B))
        bytes_seg = B * (s * d * E + s * d * Nr) + 4 * (Nr + 1)
        out.append({"kernel": "segment_sum_kernel (m2g aggregation)", "bytes": bytes_seg, "seconds": t})
    for k in out:
        k["achieved_gbs"] = k["bytes"] / k["seconds"] / 1e9
        k["frac"] = k["achieved_gbs"] / peak
    dom = max(out, key=lambda k: k["seconds"])
    # DRAM traffic per launch of the same kernel / shape from the committed ncu --set full capture
    # (dram__bytes_read.sum + dram__bytes_write.sum); only valid for the shape it was captured at
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if fused and os.path.isfile(tpath):
        tj = json.load(open(tpath)).get("by_batch", {}).get(str(B))
        if tj and (info.grid_shape if hasattr(info, "grid_shape") else None) in (None, (512, 640)) and E == 1310720:
            traffic, traffic_src = tj.get("m2g_edge_fwd_dram_bytes_per_launch"), tj.get("source")
    # `achieved` follows SURVEY 8(d): the per-sample byte count x the samples of one launch.  The edge
    # embedding is shared by the batch in this implementation and (sample-minor tile order) fetched from
    # HBM once per launch, so the measured DRAM traffic is BELOW that figure at batch > 1:
    # `achieved_dram` = measured traffic / time says how busy HBM really is.
    roof = {
        "bound": "hbm", "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s",
        "frac": dom["frac"], "traffic": traffic, "traffic_source": traffic_src, "kernel": dom["kernel"],
        "achieved_dram": (traffic / dom["seconds"] / 1e9) if traffic else None,
        "frac_dram": (traffic / dom["seconds"] / 1e9 / peak) if traffic else None,
        "peak_source": peak_src,
        "algorithmic_bytes_per_launch": dom["bytes"], "avg_launch_ms": dom["seconds"] * 1e3,
        "kernels": [{"kernel": k["kernel"], "achieved_gbs": round(k["achieved_gbs"], 1),
                     "frac": round(k["frac"], 4), "ms": round(k["seconds"] * 1e3, 4)} for k in out],
    }
    return roof


# ------------------------------------------------------------------------ GPU arm
def gpu_arm(args):
    import torch.distributed as dist

    from py4cast_b200 import _lib
    from py4cast_b200.datasets import ItemBatch, synthetic_dataset_info
    from py4cast_b200.lightning import AutoRegressiveLightning
    from py4cast_b200.trainer import DataParallelTrainer, init_distributed

    rank, local_rank, world = init_distributed("nccl")
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    dev = torch.device("cuda", local_rank)
    cfg = workload_config(args, world)
    H, W = cfg["grid"]
    info = synthetic_dataset_info(H, W)
    tmp = tempfile.mkdtemp(prefix=f"p4c_bench_r{rank}_")
    torch.manual_seed(0)
    settings = {"tmp_dir": tmp}
    if args.model != "GraphLAM":
        settings["n_max_levels"] = 3
    if args.processor_layers != 4:
        settings["processor_layers"] = args.processor_layers
    lm = AutoRegressiveLightning(
        settings_init_args=settings, dataset_info=info, num_input_steps=2,
        num_pred_steps_train=args.pred_steps, batch_size=args.batch_per_gpu, model_name=args.model,
        losses=[{"class": "WeightedLoss", "weight": 1.0, "params": {"loss": "MSELoss", "reduction": "none"}}],
        training_strategy=args.strategy, learning_rate=1e-3, min_learning_rate=3e-7, num_warmup_steps=1000,
        betas=(0.9, 0.95), precision=args.precision,
    ).to(dev)
    trainer = DataParallelTrainer(lm, total_steps=100000)
    B, T = args.batch_per_gpu, args.pred_steps
    # synthetic N(0,1) fields, generated on the GPU (the host generator is single-threaded under
    # torchrun); the pinned host batches of the end-to-end leg are copies of GPU-generated ones.  Two
    # alternate when they fit 4 GB of pinned memory per rank, otherwise one is re-sent every step.
    def gpu_batch(seed):
        g = torch.Generator(device=dev).manual_seed(seed)
        shape = info.statics.grid_shape
        names = ["batch", "timestep", "lat", "lon", "features"]
        wn, fn = info.shortnames["input_output"], info.shortnames["forcing"]
        mk = lambda steps, F: torch.randn(B, steps, *shape, F, generator=g, device=dev, dtype=torch.float32)  # noqa: E731
        from py4cast_b200.namedtensor import NamedTensor
        return ItemBatch(
            inputs=NamedTensor(mk(2, len(wn)), names, wn), forcing=NamedTensor(mk(T, len(fn)), names, fn),
            outputs=NamedTensor(mk(T, len(wn)), names, wn))

    def to_pinned(b):
        from py4cast_b200.namedtensor import NamedTensor
        cp = lambda nt: NamedTensor(nt.tensor.cpu().pin_memory(), list(nt.names), list(nt.feature_names))  # noqa: E731
        return type(b)(inputs=cp(b.inputs), forcing=cp(b.forcing), outputs=cp(b.outputs))

    dev_batch = gpu_batch(100 + rank)
    batch_bytes = sum(getattr(dev_batch, n).tensor.numel() * 4 for n in ("inputs", "forcing", "outputs"))
    n_host = 2 if 2 * batch_bytes <= (4 << 30) else 1
    host_batches = [to_pinned(gpu_batch(200 + rank + 10 * i)) for i in range(n_host)]
    if n_host == 1:
        host_batches = host_batches * 2
    torch.cuda.empty_cache()

    if args.profile_step:
        for _ in range(3):
            trainer.train_step(dev_batch)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        trainer.train_step(dev_batch)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return None

    use_graph = not args.no_graph
    # launches of OUR kernels per step, counted on an eager step (graph replays bypass the counter)
    trainer.train_step(dev_batch)
    torch.cuda.synchronize()
    _lib.reset_launch_count()
    trainer.train_step(dev_batch)
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count()
    if use_graph:
        trainer.capture(dev_batch, warmup=1)

    def step_resident():
        return trainer.replay() if use_graph else trainer.train_step(dev_batch)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    total_ms = timed(lambda i: step_resident(), args.steps)
    clocks = sampler.stop() if sampler else None
    ms_per_step = total_ms / args.steps
    value = (B * world) / (ms_per_step * 1e-3)

    # end to end: pinned host batch -> H2D -> step -> loss read back (D2H), every step
    h2d = sum(getattr(host_batches[0], n).tensor.numel() * 4 for n in ("inputs", "forcing", "outputs"))

    def step_e2e(i):
        # every step: H2D of one pinned host batch (the NEXT step's, on the copy stream, overlapping
        # this step's kernels - the input pipelining the reference gets from its pinned DataLoader)
        # and a D2H read of this step's loss
        hb = host_batches[i % 2]
        if use_graph:
            loss = trainer.replay(hb, prefetch=host_batches[(i + 1) % 2])
        else:
            gb = type(hb)(inputs=_to(hb.inputs, dev), forcing=_to(hb.forcing, dev), outputs=_to(hb.outputs, dev))
            loss = trainer.train_step(gb)
        return float(loss.item())  # D2H of the step's loss

    for i in range(2):
        step_e2e(i)
    e2e_steps = max(3, min(args.steps, 10))
    e2e_ms = timed(step_e2e, e2e_steps) / e2e_steps
    e2e = {"value": (B * world) / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms}

    result = None
    if rank == 0:
        roof = kernel_roofline(lm, info, cfg, dev)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_reference_arm(args, cfg, steps=2, warmup=1, budget_s=min(args.cpu_budget_s, 60.0))
        result = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16-true" else "f32", "data": "synthetic",
            "config": cfg, "clocks": clocks, "e2e": e2e,
            "gpu_launches": int(launches_per_step * args.steps), "gpu_launches_per_step": int(launches_per_step),
            "cuda_graph": use_graph, "roofline": roof, "cpu_baseline": cpu,
        }
    barrier()
    if world > 1:
        dist.destroy_process_group()
    return result


def _to(nt, dev):
    from py4cast_b200.namedtensor import NamedTensor

    return NamedTensor(nt.tensor.to(dev, non_blocking=True), list(nt.names), list(nt.feature_names))


def main():
    args = parse_args()
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        if rank != 0:
            return 0
        world = int(os.environ.get("WORLD_SIZE", "1"))
        cfg = workload_config(args, max(world, args.gpus))
        cpu = cpu_reference_arm(args, cfg, steps=args.steps, warmup=args.warmup, budget_s=args.cpu_budget_s)
        line = {
            "impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": cpu["timed_steps"], "warmup": 1, "ms_per_step": 1e3 / cpu["value"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": cfg, "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0
    result = gpu_arm(args)
    if result is not None:
        print(json.dumps(result))
    return 0


if __name__ == "__main__":
    sys.exit(main())
