"""Warp-specialised, recomputing backward engine (mlp2_ws_bwd.cu) against the classic tensor-core
backward (same inputs, same packed weights; the classic one reads the h1 / LayerNorm statistics the
forward saved, the pipelined one recomputes them and the forward saves nothing) and against the
plain-PyTorch fp32 oracle, through the C ABI: dense MLPs (one, two, three K-blocks, padded odd width,
residual pass-through, tiles straddling samples, several tiles per CTA), InteractionNets (uniform
in-degree with the receiver gradient summed in the epilogue, edge update with dy_b, random graphs,
batch-shared edge embeddings), deferred weight-gradient partials over several calls (trainer), bit
reproducibility, and a whole HiLAM rollout."""
import pytest
import torch

from helpers import BF16_TOL, assert_close
from test_gpu_ws_forward import _mlp, _random_graph, _uniform_graph, _weights

pytestmark = pytest.mark.gpu

HUGE = 1 << 40


@pytest.fixture()
def lib(cuda):
    from py4cast_b200 import _lib

    lib = _lib.load()
    yield lib
    lib.p4c_set_ws_bwd_min_tiles(0)  # back to the default thresholds
    lib.p4c_set_ws_min_tiles(0)


def _both(lib, fn):
    """fn() with the classic backward engine, then with the pipelined one; returns both results."""
    lib.p4c_set_ws_bwd_min_tiles(HUGE)
    n0 = lib.p4c_ws_bwd_launch_count()
    classic = fn()
    assert lib.p4c_ws_bwd_launch_count() == n0
    lib.p4c_set_ws_bwd_min_tiles(1)
    ws = fn()
    assert lib.p4c_ws_bwd_launch_count() > n0, "the pipelined backward engine did not run"
    return classic, ws


@pytest.mark.parametrize("k1,rows,batch,residual", [
    (64, 1000, 1, True),       # residual: d x += dy pass-through in the dX epilogue
    (64, 333, 3, True),        # tiles straddle samples
    (67, 777, 2, False),       # grid embedder: 64 + 3 padded columns (second K-block, 1 chunk)
    (128, 4100, 1, False),     # two K-blocks
    (128, 70000, 1, False),    # 547 tiles on 148 CTAs: both contexts wrap several times
    (64, 129, 1, False),       # one valid row in the second tile
])
def test_dense_mlp_backward_matches_classic_engine_and_oracle(lib, cuda, k1, rows, batch, residual):
    from py4cast_b200 import ops

    ref = _mlp(k1, seed=k1 + rows)
    ws = _weights(ref, True, cuda)
    meta = ops.MlpMeta(ws, (k1,), 64)
    ops.WeightPacker([meta]).pack()
    g = torch.Generator().manual_seed(rows)
    x = torch.randn(batch, rows, k1, generator=g).bfloat16().float()
    gy = torch.randn(batch, rows, 64, generator=g).bfloat16().float()

    def run():
        for w in ws:
            w.grad = None
        xp = ops.empty_rows(x.shape, torch.bfloat16, cuda, zero=True)
        xp.copy_(x.to(cuda).bfloat16())
        xp.requires_grad_(True)
        y = ops.fused_mlp(xp, ws, residual=residual, meta=meta)
        y.backward(gy.to(cuda).bfloat16())
        return [t.detach().float().cpu() for t in [y, xp.grad] + [w.grad for w in ws]]

    classic, wsr = _both(lib, run)
    xr = x.clone().requires_grad_(True)
    yr = ref(xr) + (xr if residual else 0)
    yr.backward(gy)
    refs = [yr.detach(), xr.grad] + [p.grad for p in (ref[0].weight, ref[0].bias, ref[2].weight, ref[2].bias,
                                                      ref[3].weight, ref[3].bias)]
    names = ["y", "dx", "dW1", "db1", "dW2", "db2", "dgamma", "dbeta"]
    for n, a, b, r in zip(names, wsr, classic, refs):
        assert_close(a, r, BF16_TOL, f"{n}: pipelined vs oracle")
        assert_close(a, b, 8e-3, f"{n}: pipelined vs classic")


@pytest.mark.parametrize("kind,deg,update_edges,batch,shared_edge,nr", [
    ("uniform", 4, False, 2, False, 1500),   # m2g: receiver gradient summed over groups of 4 rows in the epilogue
    ("uniform", 4, False, 3, True, 1500),    # ... with the static edge embedding shared by the batch
    ("uniform", 1, True, 1, False, 1500),    # up / down edges with edge update: dy_b, d_edge pass-through, group of 1
    ("uniform", 8, False, 1, False, 1500),
    ("uniform", 2, True, 2, False, 1500),
    ("random", 0, True, 2, False, 1500),     # g2m-like: both node gradients through the segment-sum pass
    ("random", 0, False, 1, True, 1500),
    # at least 148 row blocks: the shared edge embedding's gradient is accumulated over the samples in TMEM
    # (dx0_batch_sum; the last row block is partial, CTAs own 1 or 2 blocks)
    ("uniform", 4, False, 3, True, 5990),
    ("uniform", 4, False, 1, True, 4800),
    ("random", 0, False, 4, True, 6000),
])
def test_interaction_net_backward(lib, cuda, kind, deg, update_edges, batch, shared_edge, nr):
    from oracle import graph_ref
    from py4cast_b200 import ops
    from py4cast_b200.graph_build import EdgeSet
    from py4cast_b200.nlam import InteractionNet

    ns = 300
    ei = _uniform_graph(nr, ns, deg, 3) if kind == "uniform" else _random_graph(nr, ns, 5000 if nr < 4000 else 30001, 3)
    E = ei.shape[1]
    es = EdgeSet(ei, torch.zeros(E, 3), ns, nr)
    torch.manual_seed(7)
    net = InteractionNet(es, 64, update_edges=update_edges, aggr="sum")
    ref = graph_ref.InteractionNet(ei, 64, update_edges=update_edges, aggr="sum", num_rec=nr)
    ref.load_state_dict(net.state_dict())
    net = net.to(cuda)
    ops.WeightPacker([m.meta for m in net.modules() if hasattr(m, "meta")]).pack()
    g = torch.Generator().manual_seed(5)
    send = torch.randn(batch, ns, 64, generator=g).bfloat16().float()
    rec = torch.randn(batch, nr, 64, generator=g).bfloat16().float()
    edge = torch.randn(*((E, 64) if shared_edge else (batch, E, 64)), generator=g).bfloat16().float()
    g_rec = torch.randn(batch, nr, 64, generator=g).bfloat16().float()
    g_edge = torch.randn(batch, E, 64, generator=g).bfloat16().float()

    def run_ref():
        s, r, e = (t.clone().requires_grad_(True) for t in (send, rec, edge))
        o = ref(s, r, e.unsqueeze(0).expand(batch, -1, -1) if shared_edge else e)
        if update_edges:
            (o[0] * g_rec).sum().add((o[1] * g_edge).sum()).backward()
            return [o[0].detach(), o[1].detach(), s.grad, r.grad, e.grad]
        (o * g_rec).sum().backward()
        return [o.detach(), s.grad, r.grad, e.grad]

    def run_gpu():
        for p in net.parameters():
            p.grad = None
        s, r, e = (t.to(cuda).bfloat16().requires_grad_(True) for t in (send, rec, edge))
        o = net(s, r, e)
        if update_edges:
            torch.autograd.backward([o[0], o[1]], [g_rec.to(cuda).bfloat16(), g_edge.to(cuda).bfloat16()])
            res = [o[0], o[1], s.grad, r.grad, e.grad]
        else:
            o.backward(g_rec.to(cuda).bfloat16())
            res = [o, s.grad, r.grad, e.grad]
        res += [torch.cat([p.grad.flatten() for p in net.parameters()])]
        return [t.detach().float().cpu() for t in res]

    nb0 = lib.p4c_ws_bwd_bsum_launch_count()
    classic, wsr = _both(lib, run_gpu)
    assert (lib.p4c_ws_bwd_bsum_launch_count() - nb0 == 1) == (shared_edge and nr >= 4000), \
        "dx0_batch_sum must serve exactly the shared edge embeddings of at least 148 row blocks"
    refs = run_ref() + [torch.cat([p.grad.flatten() for p in ref.parameters()])]
    names = (["rec_out", "edge_out"] if update_edges else ["rec_out"]) + ["d_send", "d_rec", "d_edge", "d_params"]
    for n, a, b, r in zip(names, wsr, classic, refs):
        assert_close(a, r, BF16_TOL, f"{n}: pipelined vs oracle")
        assert_close(a, b, 8e-3, f"{n}: pipelined vs classic")


def test_forward_saves_nothing_for_the_recomputing_engine(lib, cuda):
    """With the pipelined backward serving a call the forward allocates neither h1 nor statistics."""
    from py4cast_b200 import ops

    ref = _mlp(64, seed=3)
    ws = _weights(ref, True, cuda)
    meta = ops.MlpMeta(ws, (64,), 64)
    ops.WeightPacker([meta]).pack()
    x = torch.randn(2, 3000, 64, device=cuda).bfloat16().requires_grad_(True)
    lib.p4c_set_ws_bwd_min_tiles(1)
    y = ops.fused_mlp(x, ws, meta=meta)
    saved = y.grad_fn.saved_tensors
    assert saved[7] is None and saved[8] is None
    lib.p4c_set_ws_bwd_min_tiles(HUGE)
    y = ops.fused_mlp(x, ws, meta=meta)
    saved = y.grad_fn.saved_tensors
    assert saved[7] is not None and saved[8] is not None


def test_pipelined_backward_is_bit_reproducible(lib, cuda):
    from py4cast_b200 import ops

    ref = _mlp(128, seed=1)
    ws = _weights(ref, True, cuda)
    meta = ops.MlpMeta(ws, (128,), 64)
    ops.WeightPacker([meta]).pack()
    x = torch.randn(2, 30000, 128, device=cuda).bfloat16()
    gy = torch.randn(2, 30000, 64, device=cuda).bfloat16()
    lib.p4c_set_ws_bwd_min_tiles(1)

    def run():
        for w in ws:
            w.grad = None
        xx = x.clone().requires_grad_(True)
        ops.fused_mlp(xx, ws, meta=meta).backward(gy)
        return [xx.grad.clone()] + [w.grad.clone() for w in ws]

    n0 = lib.p4c_ws_bwd_launch_count()
    a = run()
    for rep in range(8):  # the two MMA warps race differently from run to run: the result must not
        b = run()
        for u, v in zip(a, b):
            assert torch.equal(u, v), f"repetition {rep}"
    assert lib.p4c_ws_bwd_launch_count() - n0 == 9


def test_trainer_step_with_deferred_partials_through_the_pipelined_backward(lib, cuda, tmp_path):
    """Three-step rollout in the trainer (weight-gradient partials overwritten by the first call of an
    MLP, accumulated by the following ones, reduced once; the partial blocks beyond the pipelined
    engine's grid must read as zeros): flat gradient equal to the classic engine's."""
    from py4cast_b200.datasets import synthetic_batch
    from py4cast_b200.trainer import DataParallelTrainer, _fresh
    from test_gpu_trainer import _lm

    lm, info = _lm(tmp_path, name="HiLAM", strategy="diff_ar", precision="bf16-true")
    lm = lm.to(cuda)
    tr = DataParallelTrainer(lm, total_steps=100)
    batch = synthetic_batch(info, 2, 2, 3, seed=0, device=cuda)

    def run():
        loss = tr._forward_backward(_fresh(batch))
        torch.cuda.synchronize()
        return [loss.detach().float().cpu(), tr.flat.grad.detach().float().cpu().clone()]

    classic, wsr = _both(lib, run)
    assert_close(wsr[0], classic[0], 2e-3, "loss: pipelined vs classic")
    assert_close(wsr[1], classic[1], 1e-2, "flat gradient: pipelined vs classic")


def test_hilam_rollout_through_the_pipelined_backward(lib, cuda, tmp_path):
    """Whole HiLAM forward + backward (bf16) with every eligible MLP call on the pipelined kernels."""
    from test_gpu_models import _build

    H, W, f_in, f_out, B = 64, 64, 11, 5, 2
    model, ref = _build("HiLAM", H, W, f_in, f_out, tmp_path, processor_layers=2)
    model = model.to(cuda)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, H * W, f_in, generator=g)
    gy = torch.randn(B, H * W, f_out, generator=g)
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    yr.backward(gy)
    lib.p4c_set_ws_min_tiles(1)
    lib.p4c_set_ws_bwd_min_tiles(1)
    n0 = lib.p4c_ws_bwd_launch_count()
    xg = x.to(cuda).bfloat16().requires_grad_(True)
    y = model(xg)
    y.backward(gy.to(cuda).bfloat16())
    assert lib.p4c_ws_bwd_launch_count() - n0 >= 10
    assert_close(y.float().cpu(), yr, BF16_TOL, "y")
    assert_close(xg.grad.float().cpu(), xr.grad, BF16_TOL, "dx")
    flat_g = torch.cat([p.grad.flatten().cpu() for p in model.parameters()])
    flat_r = torch.cat([p.grad.flatten() for p in ref.parameters()])
    assert_close(flat_g, flat_r, BF16_TOL, "all parameter grads")
