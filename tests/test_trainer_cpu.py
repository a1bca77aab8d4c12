"""Host-side logic of the data-parallel trainer on CPU: flat parameter/gradient buffers, the
world-size-2 gradient exchange over gloo, sharding and the LR schedule restatement."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from py4cast_b200.trainer import FlatParams, cosine_min_lr_lambda, shard_indices


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.SiLU(), torch.nn.Linear(7, 3))


def test_flat_params_alias_parameters_and_grads():
    m = _model()
    before = [p.detach().clone() for p in m.parameters()]
    fp = FlatParams(m)
    for p, b in zip(m.parameters(), before):
        assert torch.equal(p, b)
        assert p.data_ptr() >= fp.flat.data_ptr() and p.grad.data_ptr() >= fp.grad.data_ptr()
    m(torch.randn(4, 5)).sum().backward()
    assert float(fp.grad.abs().sum()) > 0  # autograd accumulated INTO the flat buffer
    g0 = [p.grad.clone() for p in m.parameters()]
    fp.flat.mul_(2.0)  # an optimiser writing the flat buffer updates every parameter
    for p, b in zip(m.parameters(), before):
        assert torch.allclose(p, 2 * b)
    fp.zero_grad()
    assert all(float(p.grad.abs().sum()) == 0 for p in m.parameters())
    fp.enable_kernel_sinks(True)
    assert all(p._p4c_grad.data_ptr() == p.grad.data_ptr() for p in m.parameters())
    fp.enable_kernel_sinks(False)
    assert not any(hasattr(p, "_p4c_grad") for p in m.parameters())
    assert len(g0) == 4


def test_shard_indices_partition():
    for world in (1, 2, 4, 8):
        parts = [shard_indices(19, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(19))


def test_lr_schedule_matches_transformers():
    from transformers.optimization import get_cosine_with_min_lr_schedule_with_warmup

    base, min_lr, warm, total = 1e-3, 3e-7, 10, 200
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=base)
    sched = get_cosine_with_min_lr_schedule_with_warmup(
        opt, num_warmup_steps=warm, num_training_steps=total, min_lr=min_lr
    )
    for step in range(0, 220):
        ours = base * cosine_min_lr_lambda(step, warm, total, min_lr / base)
        assert abs(ours - sched.get_last_lr()[0]) <= 1e-12 + 1e-9 * base, step
        opt.step()
        sched.step()


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from py4cast_b200.trainer import DataParallelTrainer

        m = _model()
        if rank == 1:  # ranks start from different weights: construction must broadcast rank 0's
            with torch.no_grad():
                for p in m.parameters():
                    p.add_(1.0)

        class LM(torch.nn.Module):
            learning_rate, min_learning_rate, num_warmup_steps, betas = 1e-3, 1e-6, 0, (0.9, 0.999)

            def __init__(self, net):
                super().__init__()
                self.net = net

        lm = LM(m)
        tr = DataParallelTrainer.__new__(DataParallelTrainer)  # host logic only: no CUDA optimiser
        tr.lm, tr.world, tr.rank = lm, world, rank
        tr.flat = FlatParams(lm)
        dist.broadcast(tr.flat.flat, src=0)
        g = torch.Generator().manual_seed(1)
        x, y = torch.randn(8, 5, generator=g), torch.randn(8, 3, generator=g)
        idx = shard_indices(8, rank, world)
        tr.flat.zero_grad()
        torch.nn.functional.mse_loss(m(x[idx]), y[idx]).backward()
        scale = tr.sync_gradients()
        ret[rank] = (tr.flat.flat.clone(), tr.flat.grad.clone() * scale)
    finally:
        dist.destroy_process_group()


def test_world_size_2_gradient_exchange_gloo():
    world, port = 2, 29611
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    (w0, g0), (w1, g1) = ret[0], ret[1]
    assert torch.equal(w0, w1), "weights must be identical on every rank"
    assert torch.equal(g0, g1), "all-reduced gradients must be identical on every rank"
    # equals the single-process gradient on the full batch (mean over shards of equal size)
    m = _model()
    fp = FlatParams(m)
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(8, 5, generator=g), torch.randn(8, 3, generator=g)
    torch.nn.functional.mse_loss(m(x), y).backward()
    assert torch.allclose(fp.grad, g0, rtol=1e-5, atol=1e-7)
