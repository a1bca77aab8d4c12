"""Tensor-core engine features (bf16 storage), through the C ABI: padded odd-width sources, packed
weight images, segment-sum prologue, aggregation in the epilogue, fused segment-sum + GEMM, and the
deferred (batched) weight-gradient reduction.  References: plain PyTorch fp32 of the same op."""
import ctypes as C

import pytest
import torch

from helpers import BF16_TOL, assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(cuda):
    from py4cast_b200 import ops as _ops

    return _ops


def _ref_mlp(k1, out, ln, seed):
    from oracle.graph_ref import make_mlp

    torch.manual_seed(seed)
    m = make_mlp([k1, 64, out], layer_norm=ln)
    if ln:
        with torch.no_grad():
            m[3].weight.uniform_(0.5, 1.5)
            m[3].bias.normal_(0, 0.1)
    return m


def _weights(m, ln, dev, grad=True):
    ws = [m[0].weight, m[0].bias, m[2].weight, m[2].bias]
    ws += [m[3].weight, m[3].bias] if ln else [None, None]
    return [w.detach().to(dev).requires_grad_(grad) if w is not None else None for w in ws]


def _padded(ops, x, dev):
    """bf16 copy of x whose rows are padded to 16 bytes (zeros), as the (…, W) view."""
    p = ops.empty_rows(x.shape, torch.bfloat16, dev, zero=True)
    p.copy_(x.to(dev).bfloat16())
    return p


def _packed_meta(ops, ws, widths, out):
    meta = ops.MlpMeta(ws, widths, out)
    ops.WeightPacker([meta]).pack()
    assert meta.image is not None
    return meta


@pytest.mark.parametrize("k1,out,ln,rows,batch", [
    (67, 64, True, 777, 2), (3, 64, True, 130, 1), (2, 64, True, 81, 1), (128, 64, True, 300, 1),
    (100, 64, True, 515, 1), (72, 21, False, 260, 2), (192, 64, True, 129, 1),
])
@pytest.mark.parametrize("packed", [False, True])
def test_padded_width_sources_forward_backward(ops, cuda, k1, out, ln, rows, batch, packed):
    """Dense sources of any width ride the tensor cores once their rows are 16-byte padded."""
    from py4cast_b200 import _lib

    ref = _ref_mlp(k1, out, ln, seed=k1 + out)
    ws = _weights(ref, ln, cuda)
    g = torch.Generator().manual_seed(rows)
    x = torch.randn(batch, rows, k1, generator=g).bfloat16().float()
    gy = torch.randn(batch, rows, out, generator=g).bfloat16().float()
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    yr.backward(gy)
    xg = _padded(ops, x, cuda).requires_grad_(True)
    assert xg.stride(-2) % 8 == 0
    meta = _packed_meta(ops, ws, (k1,), out) if packed else None
    lib = _lib.load()
    n0 = lib.p4c_launch_count()
    y = ops.fused_mlp(xg, ws, meta=meta)
    y.backward(gy.to(cuda).bfloat16())
    assert lib.p4c_launch_count() - n0 == 3  # forward, backward, weight-gradient reduction
    assert_close(y.float().cpu(), yr, BF16_TOL, "y")
    assert_close(xg.grad.float().cpu(), xr.grad, BF16_TOL, "dx")
    refs = [ref[0].weight, ref[0].bias, ref[2].weight, ref[2].bias] + ([ref[3].weight, ref[3].bias] if ln else [])
    for n, w, r in zip(["w1", "b1", "w2", "b2", "gamma", "beta"], ws, refs):
        assert_close(w.grad.cpu(), r.grad, BF16_TOL, f"d{n}")


def test_packed_image_is_bit_identical_to_in_kernel_conversion(ops, cuda):
    ref = _ref_mlp(192, 64, True, seed=4)
    ws = _weights(ref, True, cuda)
    g = torch.Generator().manual_seed(0)
    srcs = [torch.randn(1, 3000, 64, generator=g).to(cuda).bfloat16() for _ in range(3)]
    x = torch.cat(srcs, -1)  # (1, 3000, 192): one 192-wide source = 3 K-blocks
    outs = []
    for packed in (False, True):
        for w in ws:
            w.grad = None
        xg = x.clone().requires_grad_(True)
        meta = _packed_meta(ops, ws, (192,), 64) if packed else None
        y = ops.fused_mlp(xg, ws, meta=meta)
        y.float().square().sum().backward()
        outs.append([y.detach(), xg.grad] + [w.grad.clone() for w in ws])
    for a, b in zip(*outs):
        assert torch.equal(a, b)


def _uniform_graph(nr, ns, deg, seed):
    g = torch.Generator().manual_seed(seed)
    rec = torch.arange(nr).repeat_interleave(deg)
    send = torch.randint(0, ns, (nr * deg,), generator=g)
    send[0], send[1] = ns - 1, 0
    return torch.stack([send, rec])


@pytest.mark.parametrize("deg", [1, 4, 8])
@pytest.mark.parametrize("aggr", ["sum", "mean"])
@pytest.mark.parametrize("update_edges", [False, True])
def test_interaction_net_uniform_degree_epilogue_aggregation(ops, cuda, deg, aggr, update_edges):
    """Receiver-sorted edges with a uniform in-degree (m2g: 4, up/down: 1): the aggregation runs in
    the epilogue of the edge kernel and the messages are never written."""
    from oracle import graph_ref
    from py4cast_b200 import _lib
    from py4cast_b200.graph_build import EdgeSet
    from py4cast_b200.nlam import InteractionNet

    nr, ns, B = 517, 90, 2
    ei = _uniform_graph(nr, ns, deg, seed=deg)
    E = ei.shape[1]
    torch.manual_seed(5)
    ref = graph_ref.InteractionNet(ei, 64, update_edges=update_edges, aggr=aggr)
    net = InteractionNet(EdgeSet(ei, torch.zeros(E, 3), ns, nr), 64, update_edges=update_edges, aggr=aggr)
    net.load_state_dict(ref.state_dict())
    net = net.to(cuda)
    assert net.graph_index().uniform_deg == deg
    g = torch.Generator().manual_seed(6)
    send, rec, edge = (torch.randn(B, n, 64, generator=g).bfloat16().float() for n in (ns, nr, E))
    g_rec, g_edge = torch.randn(B, nr, 64, generator=g), torch.randn(B, E, 64, generator=g)
    ins_r = [t.clone().requires_grad_(True) for t in (send, rec, edge)]
    out_r = ref(*ins_r)
    ins_g = [t.to(cuda).bfloat16().requires_grad_(True) for t in (send, rec, edge)]
    lib = _lib.load()
    n0 = lib.p4c_launch_count()
    out_g = net(*ins_g)
    assert lib.p4c_launch_count() - n0 == 2, "forward = edge kernel (+aggregation) and node kernel"
    if update_edges:
        (out_r[0] * g_rec).sum().add((out_r[1] * g_edge).sum()).backward()
        (out_g[0].float() * g_rec.to(cuda)).sum().add((out_g[1].float() * g_edge.to(cuda)).sum()).backward()
        assert_close(out_g[1].float().cpu(), out_r[1], BF16_TOL, "edge_out")
        assert_close(out_g[0].float().cpu(), out_r[0], BF16_TOL, "rec_out")
    else:
        (out_r * g_rec).sum().backward()
        (out_g.float() * g_rec.to(cuda)).sum().backward()
        assert_close(out_g.float().cpu(), out_r, BF16_TOL, "rec_out")
    for n, a, b in zip(("send", "rec", "edge"), ins_g, ins_r):
        assert_close(a.grad.float().cpu(), b.grad, BF16_TOL, f"d_{n}")
    for (n, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        assert_close(p.grad.cpu(), q.grad, BF16_TOL, f"d_{n}")


def test_interaction_net_launch_counts_general_graph(ops, cuda):
    """General (ragged, unsorted) edge set in bf16: 2 launches forward (aggregation in the prologue
    of the node kernel), 3 backward (node MLP, edge MLP, ONE paired segment-sum + GEMM launch for the sender
    and the receiver gradient) + 2 weight-gradient reductions when nothing is deferred."""
    from py4cast_b200 import _lib
    from py4cast_b200.graph_build import EdgeSet
    from py4cast_b200.nlam import InteractionNet

    g = torch.Generator().manual_seed(1)
    E, ns, nr, B = 5000, 300, 200, 1
    ei = torch.stack([torch.randint(0, ns, (E,), generator=g), torch.randint(0, nr, (E,), generator=g)])
    net = InteractionNet(EdgeSet(ei, torch.zeros(E, 3), ns, nr), 64).to(cuda)
    assert net.graph_index().uniform_deg == 0
    ins = [torch.randn(B, n, 64, generator=g).to(cuda).bfloat16().requires_grad_(True) for n in (ns, nr, E)]
    lib = _lib.load()
    n0 = lib.p4c_launch_count()
    rec_out, edge_out = net(*ins)
    assert lib.p4c_launch_count() - n0 == 2
    n0 = lib.p4c_launch_count()
    (rec_out.float().sum() + edge_out.float().sum()).backward()
    assert lib.p4c_launch_count() - n0 == 5


def test_segment_sum_prologue_and_seg_linear_vs_index_add(ops, cuda):
    """p4c_seg_linear == (index_add of the rows) @ W slice; run-to-run bit-reproducible."""
    g = torch.Generator().manual_seed(3)
    E, N, B = 20000, 900, 2
    keys = torch.randint(0, N - 50, (E,), generator=g, dtype=torch.int64).to(cuda)  # some empty segments
    src = torch.randn(B, E, 64, generator=g).to(cuda).bfloat16()
    w1 = (torch.randn(64, 192, generator=g) * 0.1).to(cuda)
    csr = ops.csr_build(keys, N)
    summed = torch.zeros(B, N, 64, device=cuda).index_add_(1, keys, src.float())
    for col0 in (0, 64, 128):
        want = summed @ w1[:, col0:col0 + 64]
        out = torch.zeros(B, N, 64, device=cuda, dtype=torch.bfloat16)
        ops.seg_linear(src, csr, B, w1, col0, out, accumulate=False)
        assert_close(out.float(), want, BF16_TOL, f"seg_linear col0={col0}")
        out2 = torch.zeros_like(out)
        ops.seg_linear(src, csr, B, w1, col0, out2, accumulate=False)
        assert torch.equal(out, out2)
        ops.seg_linear(src, csr, B, w1, col0, out2, accumulate=True)
        assert_close(out2.float(), 2 * want, BF16_TOL, "accumulate")
    # mean scaling
    deg = torch.bincount(keys, minlength=N).clamp(min=1).view(1, N, 1)
    out = torch.zeros(B, N, 64, device=cuda, dtype=torch.bfloat16)
    ops.seg_linear(src, csr, B, w1, 64, out, accumulate=False, mean=True)
    assert_close(out.float(), (summed / deg) @ w1[:, 64:128], BF16_TOL, "mean")


def test_seg_linear_pair_matches_two_calls(ops, cuda):
    """p4c_seg_linear_pair: two segment sets over the same rows in one launch -- independent outputs (CTAs
    split) are bit-equal to two p4c_seg_linear calls; a shared output (senders are receivers) is ONE product
    over two K-blocks with a single update of `out` (vs index_add reference, and bit-reproducible)."""
    g = torch.Generator().manual_seed(5)
    for E, N, M, B in ((20000, 900, 700, 2), (700, 81, 81, 16), (51520, 6561, 6561, 3)):
        k_rec = torch.randint(0, N, (E,), generator=g, dtype=torch.int64).to(cuda)
        k_send = torch.randint(0, M, (E,), generator=g, dtype=torch.int64).to(cuda)
        src = torch.randn(B, E, 64, generator=g).to(cuda).bfloat16()
        w1 = (torch.randn(64, 192, generator=g) * 0.1).to(cuda)
        c_rec, c_send = ops.csr_build(k_rec, N), ops.csr_build(k_send, M)
        base_r = torch.randn(B, N, 64, generator=g).to(cuda).bfloat16()
        # independent outputs
        o_r, o_s = base_r.clone(), torch.full((B, M, 64), 7.0, device=cuda, dtype=torch.bfloat16)
        ops.seg_linear_pair(src, w1, None, B, (c_rec, 128, o_r, True), (c_send, 64, o_s, False))
        w_r, w_s = base_r.clone(), torch.empty_like(o_s)
        ops.seg_linear(src, c_rec, B, w1, 128, w_r, accumulate=True)
        ops.seg_linear(src, c_send, B, w1, 64, w_s, accumulate=False)
        assert torch.equal(o_r, w_r) and torch.equal(o_s, w_s), (E, N, M, B)
        if N != M:
            continue
        # shared output: out = base + S_rec W_r + S_send W_s, one rounding
        s_r = torch.zeros(B, N, 64, device=cuda).index_add_(1, k_rec, src.float())
        s_s = torch.zeros(B, N, 64, device=cuda).index_add_(1, k_send, src.float())
        want = base_r.float() + s_r @ w1[:, 128:192] + s_s @ w1[:, 64:128]
        o1, o2 = base_r.clone(), base_r.clone()
        ops.seg_linear_pair(src, w1, None, B, (c_rec, 128, o1, True), (c_send, 64, o1, True))
        ops.seg_linear_pair(src, w1, None, B, (c_rec, 128, o2, True), (c_send, 64, o2, True))
        assert_close(o1.float(), want, BF16_TOL, f"shared output E={E}")
        assert torch.equal(o1, o2)


def test_gradient_hand_over_between_two_consumers(ops, cuda):
    """share_grads: a representation that feeds an MLP and, as its senders, an InteractionNet gets ONE gradient
    tensor (the by-sender kernel adds into the MLP's input gradient) -- same numbers as autograd's own sum, also
    on repeated backward passes over the same storage (entries of the registry die with the autograd pass)."""
    from py4cast_b200.graph_build import EdgeSet
    from py4cast_b200.nlam import InteractionNet, make_mlp

    g = torch.Generator().manual_seed(11)
    E, ns, nr, B = 6000, 400, 250, 2
    ei = torch.stack([torch.randint(0, ns, (E,), generator=g), torch.randint(0, nr, (E,), generator=g)])
    torch.manual_seed(3)
    net = InteractionNet(EdgeSet(ei, torch.zeros(E, 3), ns, nr), 64, update_edges=False).to(cuda)
    mlp = make_mlp([64, 64, 64]).to(cuda)
    x = torch.randn(B, ns, 64, generator=g).to(cuda).bfloat16()
    rec = torch.randn(B, nr, 64, generator=g).to(cuda).bfloat16()
    edge = torch.randn(B, E, 64, generator=g).to(cuda).bfloat16()
    gy = torch.randn(B, nr, 64, generator=g).to(cuda).bfloat16()
    gm = torch.randn(B, ns, 64, generator=g).to(cuda).bfloat16()

    def run(share):
        xs = x.clone().requires_grad_(True)
        h = xs * 1.0  # a produced tensor: two consumers below
        out = net(h, rec, edge, share_grads=share)
        y = mlp(h, residual=True, share_grads=share)
        torch.autograd.backward([out, y], [gy, gm])
        return xs.grad.float()

    want = run(False)
    for _ in range(3):  # recycled addresses across passes
        got = run(True)
        assert_close(got, want, 8e-3, "handed-over gradient vs autograd's sum")


def test_deferred_weight_gradient_reduction_matches_immediate(ops, cuda):
    """Trainer mode: parameters with gradient sinks keep per-CTA partials across calls and are
    reduced by ONE batched launch (flush_wgrads); same numbers as the in-call reduction."""
    from py4cast_b200 import _lib

    lib = _lib.load()
    ref = _ref_mlp(64, 64, True, seed=9)
    g = torch.Generator().manual_seed(2)
    xs = [torch.randn(2, 3000, 64, generator=g).to(cuda).bfloat16() for _ in range(3)]
    gys = [torch.randn(2, 3000, 64, generator=g).to(cuda).bfloat16() for _ in range(3)]
    # immediate
    ws = _weights(ref, True, cuda)
    for x, gy in zip(xs, gys):
        ops.fused_mlp(x.clone().requires_grad_(True), ws).backward(gy)
    want = [w.grad.clone() for w in ws]
    # deferred: sinks + meta
    ws2 = _weights(ref, True, cuda)
    sinks = [torch.zeros_like(w) for w in ws2]
    for w, s in zip(ws2, sinks):
        w._p4c_grad = s
    meta = _packed_meta(ops, ws2, (64,), 64)
    n0 = lib.p4c_launch_count()
    for x, gy in zip(xs, gys):
        ops.fused_mlp(x.clone().requires_grad_(True), ws2, meta=meta).backward(gy)
    assert lib.p4c_launch_count() - n0 == 6, "3 x (forward + backward), no per-call reduction"
    assert all(float(s.abs().sum()) == 0.0 for s in sinks)
    ops.flush_wgrads()
    assert lib.p4c_launch_count() - n0 == 7
    for n, s, w in zip(["w1", "b1", "w2", "b2", "gamma", "beta"], sinks, want):
        assert_close(s, w, 1e-5, f"deferred d{n}")
    # a second round starts from fresh partials (and accumulates into the sinks like autograd would)
    for x, gy in zip(xs, gys):
        ops.fused_mlp(x.clone().requires_grad_(True), ws2, meta=meta).backward(gy)
    ops.flush_wgrads()
    for n, s, w in zip(["w1", "b1", "w2", "b2", "gamma", "beta"], sinks, want):
        assert_close(s, 2 * w, 1e-5, f"second flush d{n}")


def test_forward_aggregate_rejects_bad_groups(ops, cuda):
    from py4cast_b200 import _lib
    from py4cast_b200._lib import Mlp

    lib = _lib.load()
    ref = _ref_mlp(64, 64, True, seed=1)
    ws = _weights(ref, True, cuda, grad=False)
    x = torch.randn(1, 130, 64).to(cuda).bfloat16()
    aggr = torch.zeros(1, 65, 64, device=cuda, dtype=torch.bfloat16)
    desc = Mlp()
    ops._fill_mlp(desc, 130, 1, 0, [ops._SrcSpec(x)], ws, 64, torch.bfloat16)
    for group in (0, 3, 16, 4):  # 4 does not divide 130
        rc = lib.p4c_mlp_forward_aggregate(C.byref(desc), None, None, None, None, _lib.ptr(aggr), group, 1.0,
                                           _lib.stream())
        assert rc == -1, group
    rc = lib.p4c_mlp_forward_aggregate(C.byref(desc), None, None, None, None, _lib.ptr(aggr), 2, 1.0, _lib.stream())
    assert rc == 0
    torch.cuda.synchronize()

