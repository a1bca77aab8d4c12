"""Parity of the individual CUDA kernels (through the C ABI) against plain-PyTorch fp32
references / the oracle: CSR, segment sum, fused MLP, InteractionNet."""
import pytest
import torch

from helpers import BF16_TOL, FP32_TOL, assert_close, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(cuda):
    from py4cast_b200 import ops as _ops

    return _ops


# ----------------------------------------------------------------------------- CSR
@pytest.mark.parametrize("E,K", [(0, 5), (1, 1), (1000, 7), (5436, 81), (200000, 6561), (50000, 50000)])
def test_csr_build_bit_exact_vs_stable_sort(ops, cuda, E, K):
    g = torch.Generator().manual_seed(E + K)
    keys = torch.randint(0, K, (E,), generator=g, dtype=torch.int64).to(cuda)
    csr = ops.csr_build(keys, K, drop_identity_perm=False)
    order = torch.sort(keys, stable=True).indices.to(torch.int32)
    counts = torch.bincount(keys, minlength=K)
    rowptr = torch.cat([torch.zeros(1, dtype=torch.int64, device=cuda), counts.cumsum(0)]).to(torch.int32)
    assert torch.equal(csr.rowptr, rowptr)
    assert torch.equal(csr.perm, order)


def test_csr_sorted_input_gives_identity(ops, cuda):
    keys = torch.arange(1000, device=cuda).repeat_interleave(4)
    csr = ops.csr_build(keys, 1000)
    assert csr.perm is None and int(csr.rowptr[-1]) == 4000


def test_csr_rejects_bad_arguments(ops, cuda):
    with pytest.raises(TypeError):
        ops.csr_build(torch.zeros(4, dtype=torch.int32, device=cuda), 2)


# --------------------------------------------------------------------- segment sum
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("mean", [False, True])
def test_segment_sum_matches_index_add(ops, cuda, dtype, mean):
    g = torch.Generator().manual_seed(3)
    E, N, B = 20000, 900, 2
    keys = torch.randint(0, N - 50, (E,), generator=g, dtype=torch.int64).to(cuda)  # some empty segments
    src = torch.randn(B, E, 64, generator=g).to(cuda).to(dtype)
    csr = ops.csr_build(keys, N)
    out = ops.segment_sum(src, csr, B, mean=mean)
    ref = torch.zeros(B, N, 64, device=cuda).index_add_(1, keys, src.float())
    if mean:
        ref = ref / torch.bincount(keys, minlength=N).clamp(min=1).view(1, N, 1)
    assert_close(out.float(), ref, FP32_TOL if dtype == torch.float32 else BF16_TOL)
    # bit-reproducible run to run
    assert torch.equal(out, ops.segment_sum(src, csr, B, mean=mean))
    # accumulate
    acc = ops.segment_sum(src, csr, B, mean=mean, out=out.clone(), accumulate=True)
    assert_close(acc.float(), 2 * ref, FP32_TOL if dtype == torch.float32 else BF16_TOL)


def test_segment_sum_linearity(ops, cuda):
    g = torch.Generator().manual_seed(4)
    keys = torch.sort(torch.randint(0, 300, (5000,), generator=g)).values.to(cuda)
    a = torch.randn(1, 5000, 64, generator=g).to(cuda)
    b = torch.randn(1, 5000, 64, generator=g).to(cuda)
    csr = ops.csr_build(keys, 300)
    assert csr.perm is None
    lhs = ops.segment_sum(a + b, csr, 1)
    rhs = ops.segment_sum(a, csr, 1) + ops.segment_sum(b, csr, 1)
    assert_close(lhs, rhs, 1e-6)


# ----------------------------------------------------------------------- fused MLP
def _ref_mlp(k1, out, ln, seed):
    from oracle.graph_ref import make_mlp

    torch.manual_seed(seed)
    return make_mlp([k1, 64, out], layer_norm=ln)


def _weights(m, ln, dev):
    ws = [m[0].weight, m[0].bias, m[2].weight, m[2].bias]
    ws += [m[3].weight, m[3].bias] if ln else [None, None]
    return [w.detach().to(dev).requires_grad_(True) if w is not None else None for w in ws]


@pytest.mark.parametrize("k1,out,ln,rows,batch", [
    (64, 64, True, 1000, 2), (67, 64, True, 777, 1), (3, 64, True, 130, 1), (2, 64, True, 81, 1),
    (64, 21, False, 500, 2), (64, 1, False, 129, 1), (128, 64, True, 300, 1),
])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fused_mlp_forward_backward(ops, cuda, k1, out, ln, rows, batch, dtype):
    ref = _ref_mlp(k1, out, ln, seed=k1 + out)
    if ln:  # non-trivial affine
        with torch.no_grad():
            ref[3].weight.uniform_(0.5, 1.5)
            ref[3].bias.normal_(0, 0.1)
    ws = _weights(ref, ln, cuda)
    g = torch.Generator().manual_seed(rows)
    x = torch.randn(batch, rows, k1, generator=g)
    gy = torch.randn(batch, rows, out, generator=g)
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    yr.backward(gy)
    xg = x.to(cuda).to(dtype).requires_grad_(True)
    y = ops.fused_mlp(xg, ws)
    y.backward(gy.to(cuda).to(dtype))
    tol = FP32_TOL if dtype == torch.float32 else BF16_TOL
    assert y.dtype == dtype
    assert_close(y.float().cpu(), yr, tol, "y")
    assert_close(xg.grad.float().cpu(), xr.grad, tol, "dx")
    names = ["w1", "b1", "w2", "b2", "gamma", "beta"]
    refs = [ref[0].weight, ref[0].bias, ref[2].weight, ref[2].bias] + ([ref[3].weight, ref[3].bias] if ln else [])
    for n, w, r in zip(names, ws, refs):
        assert_close(w.grad.cpu(), r.grad, tol, f"d{n}")


def test_fused_mlp_residual_and_shared_input(ops, cuda):
    ref = _ref_mlp(64, 64, True, seed=1)
    ws = _weights(ref, True, cuda)
    x = torch.randn(2, 333, 64)
    xr = x.clone().requires_grad_(True)
    yr = xr + ref(xr)
    yr.sum().backward()
    xg = x.to(cuda).requires_grad_(True)
    y = ops.fused_mlp(xg, ws, residual=True)
    y.sum().backward()
    assert_close(y.cpu(), yr, FP32_TOL)
    assert_close(xg.grad.cpu(), xr.grad, FP32_TOL)
    # 2-D (batch-shared) input
    y2 = ops.fused_mlp(x[0].to(cuda), [w.detach() for w in ws])
    assert y2.shape == (333, 64)
    assert_close(y2.cpu(), ref(x[0]), FP32_TOL)


def test_fused_mlp_is_bit_reproducible(ops, cuda):
    ref = _ref_mlp(64, 64, True, seed=2)
    ws = _weights(ref, True, cuda)
    x = torch.randn(1, 5000, 64).to(cuda).requires_grad_(True)
    outs = []
    for _ in range(2):
        for w in ws:
            w.grad = None
        x.grad = None
        y = ops.fused_mlp(x, ws)
        y.square().sum().backward()
        outs.append([y.detach().clone(), x.grad.clone()] + [w.grad.clone() for w in ws])
    for a, b in zip(*outs):
        assert torch.equal(a, b)


# ------------------------------------------------------------------ InteractionNet
def _random_graph(E, ns, nr, sorted_by_rec, seed):
    g = torch.Generator().manual_seed(seed)
    send = torch.randint(0, ns, (E,), generator=g)
    rec = torch.randint(0, nr, (E,), generator=g)
    send[0], rec[0] = ns - 1, nr - 1  # every index range is hit (oracle re-bases by min / max)
    send[1], rec[1] = 0, 0
    if sorted_by_rec:
        order = torch.sort(rec * ns + send, stable=True).indices
        send, rec = send[order], rec[order]
    return torch.stack([send, rec])


@pytest.mark.parametrize("update_edges", [True, False])
@pytest.mark.parametrize("aggr", ["sum", "mean"])
@pytest.mark.parametrize("sorted_by_rec", [True, False])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_interaction_net_vs_oracle(ops, cuda, update_edges, aggr, sorted_by_rec, dtype):
    from oracle import graph_ref
    from py4cast_b200.graph_build import EdgeSet
    from py4cast_b200.nlam import InteractionNet

    E, ns, nr, B = 3000, 400, 150, 2
    ei = _random_graph(E, ns, nr, sorted_by_rec, seed=11)
    torch.manual_seed(5)
    ref = graph_ref.InteractionNet(ei, 64, update_edges=update_edges, aggr=aggr)
    net = InteractionNet(EdgeSet(ei, torch.zeros(E, 3), ns, nr), 64, update_edges=update_edges, aggr=aggr)
    net.load_state_dict(ref.state_dict())
    net = net.to(cuda)
    g = torch.Generator().manual_seed(6)
    send, rec, edge = (torch.randn(B, n, 64, generator=g) for n in (ns, nr, E))
    g_rec, g_edge = torch.randn(B, nr, 64, generator=g), torch.randn(B, E, 64, generator=g)
    ins_r = [t.clone().requires_grad_(True) for t in (send, rec, edge)]
    out_r = ref(*ins_r)
    ins_g = [t.to(cuda).to(dtype).requires_grad_(True) for t in (send, rec, edge)]
    out_g = net(*ins_g)
    tol = FP32_TOL if dtype == torch.float32 else BF16_TOL
    if update_edges:
        (out_r[0] * g_rec).sum().add((out_r[1] * g_edge).sum()).backward()
        (out_g[0].float() * g_rec.to(cuda)).sum().add((out_g[1].float() * g_edge.to(cuda)).sum()).backward()
        assert_close(out_g[1].float().cpu(), out_r[1], tol, "edge_out")
        assert_close(out_g[0].float().cpu(), out_r[0], tol, "rec_out")
    else:
        (out_r * g_rec).sum().backward()
        (out_g.float() * g_rec.to(cuda)).sum().backward()
        assert_close(out_g.float().cpu(), out_r, tol, "rec_out")
    for n, a, b in zip(("send", "rec", "edge"), ins_g, ins_r):
        assert_close(a.grad.float().cpu(), b.grad, tol, f"d_{n}")
    for (n, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        assert_close(p.grad.cpu(), q.grad, tol, f"d_{n}")


def test_interaction_net_shared_inputs_and_same_node_set(ops, cuda):
    """mesh 'same' GNN: send is rec; g2m: receiver + edge embeddings shared across the batch."""
    from oracle import graph_ref
    from py4cast_b200.graph_build import EdgeSet
    from py4cast_b200.nlam import InteractionNet

    E, n, B = 2000, 120, 3
    ei = _random_graph(E, n, n, True, seed=21)
    torch.manual_seed(8)
    ref = graph_ref.InteractionNet(ei, 64, update_edges=False)
    net = InteractionNet(EdgeSet(ei, torch.zeros(E, 3), n, n), 64, update_edges=False)
    net.load_state_dict(ref.state_dict())
    net = net.to(cuda)
    send = torch.randn(B, n, 64)
    rec2d, edge2d = torch.randn(n, 64), torch.randn(E, 64)
    r_in = [send.clone().requires_grad_(True), rec2d.clone().requires_grad_(True), edge2d.clone().requires_grad_(True)]
    out_r = ref(r_in[0], r_in[1].unsqueeze(0).expand(B, n, 64), r_in[2].unsqueeze(0).expand(B, E, 64))
    out_r.square().sum().backward()
    g_in = [t.detach().to(cuda).requires_grad_(True) for t in (send, rec2d, edge2d)]
    out_g = net(*g_in)
    out_g.square().sum().backward()
    assert_close(out_g.cpu(), out_r, FP32_TOL)
    for a, b in zip(g_in, r_in):
        assert a.grad.shape == b.grad.shape
        assert_close(a.grad.cpu(), b.grad, FP32_TOL)
    # send is rec (same tensor object feeds both roles)
    ref2 = graph_ref.InteractionNet(ei, 64)
    net2 = InteractionNet(EdgeSet(ei, torch.zeros(E, 3), n, n), 64)
    net2.load_state_dict(ref2.state_dict())
    net2 = net2.to(cuda)
    x, e = torch.randn(B, n, 64), torch.randn(B, E, 64)
    xr, er = x.clone().requires_grad_(True), e.clone().requires_grad_(True)
    a, b = ref2(xr, xr, er)
    (a.sum() + b.square().sum()).backward()
    xg, eg = x.to(cuda).requires_grad_(True), e.to(cuda).requires_grad_(True)
    c, d = net2(xg, xg, eg)
    (c.sum() + d.square().sum()).backward()
    assert_close(xg.grad.cpu(), xr.grad, FP32_TOL)
    assert_close(eg.grad.cpu(), er.grad, FP32_TOL)


def test_interaction_net_chunked_weights(ops, cuda):
    """HiLAMParallel-style SplitMLPs: per-edge-chunk and per-node-chunk weights."""
    from oracle import graph_ref
    from py4cast_b200.graph_build import EdgeSet
    from py4cast_b200.nlam import InteractionNet

    n, B = 200, 2
    chunks_e, chunks_n = [700, 300, 500], [120, 80]
    E = sum(chunks_e)
    ei = _random_graph(E, n, n, False, seed=31)
    torch.manual_seed(9)
    ref = graph_ref.InteractionNet(ei, 64, edge_chunk_sizes=chunks_e, aggr_chunk_sizes=chunks_n)
    net = InteractionNet(EdgeSet(ei, torch.zeros(E, 3), n, n), 64, edge_chunk_sizes=chunks_e,
                         aggr_chunk_sizes=chunks_n)
    net.load_state_dict(ref.state_dict())
    net = net.to(cuda)
    x, e = torch.randn(B, n, 64), torch.randn(B, E, 64)
    xr, er = x.clone().requires_grad_(True), e.clone().requires_grad_(True)
    a, b = ref(xr, xr, er)
    (a.square().sum() + b.square().sum()).backward()
    xg, eg = x.to(cuda).requires_grad_(True), e.to(cuda).requires_grad_(True)
    c, d = net(xg, xg, eg)
    (c.square().sum() + d.square().sum()).backward()
    assert_close(c.cpu(), a, FP32_TOL)
    assert_close(d.cpu(), b, FP32_TOL)
    assert_close(xg.grad.cpu(), xr.grad, FP32_TOL)
    assert_close(eg.grad.cpu(), er.grad, FP32_TOL)
    for (nm, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        assert_close(p.grad.cpu(), q.grad, FP32_TOL, nm)


# ------------------------------------------------------- tensor-core engine vs CUDA-core engine
@pytest.mark.parametrize("nsrc,out,ln,rows,batch", [
    (1, 64, True, 1000, 2), (2, 64, True, 300, 1), (3, 64, True, 4097, 1), (1, 21, False, 513, 3),
    (1, 64, True, 40000, 1),  # > 2 x 148 tiles: persistent loop + prefetch path
])
def test_tcgen05_engine_matches_simt_engine(ops, cuda, nsrc, out, ln, rows, batch):
    """Same call through both engines (bf16 storage): forward outputs and saved activations."""
    import ctypes as C

    from py4cast_b200 import _lib
    from py4cast_b200._lib import Mlp, check, ptr, stream

    lib = _lib.load()
    ref = _ref_mlp(64 * nsrc, out, ln, seed=nsrc)
    ws = [w.detach() if w is not None else None for w in _weights(ref, ln, cuda)]
    g = torch.Generator().manual_seed(rows)
    srcs = [torch.randn(batch, rows, 64, generator=g).to(cuda).bfloat16() for _ in range(nsrc)]
    resid = torch.randn(batch, rows, out, generator=g).to(cuda).bfloat16()
    res = {}
    for engine in (0, 1):
        lib.p4c_set_engine(engine)
        o = torch.zeros(batch, rows, out, device=cuda, dtype=torch.bfloat16)
        o2 = torch.zeros(batch, rows, 64, device=cuda, dtype=torch.bfloat16) if out == 64 else None
        h1 = torch.zeros(batch, rows, 64, device=cuda, dtype=torch.bfloat16)
        st = torch.zeros(batch, rows, 2, device=cuda)
        desc = Mlp()
        ops._fill_mlp(desc, rows, batch, 0, [ops._SrcSpec(s) for s in srcs], ws, out, torch.bfloat16)
        check(lib.p4c_mlp_forward(C.byref(desc), ptr(resid), ptr(o), ptr(o2), ptr(h1), ptr(st) if ln else None,
                                  stream()), "fwd")
        torch.cuda.synchronize()
        res[engine] = (o.float(), o2.float() if o2 is not None else None, h1.float(), st)
    lib.p4c_set_engine(1)
    # fp32 reference on the same bf16 inputs
    x = torch.cat([s.float() for s in srcs], -1).cpu()
    want = ref(x) + resid.float().cpu()
    assert_close(res[1][0].cpu(), want, BF16_TOL, "tc out vs fp32 reference")
    assert_close(res[1][0], res[0][0], BF16_TOL, "tc out vs simt out")
    assert_close(res[1][2], res[0][2], BF16_TOL, "tc h1 vs simt h1")
    if out == 64:
        assert_close(res[1][1], res[0][1], BF16_TOL, "tc out2 vs simt out2")
    if ln:
        assert_close(res[1][3], res[0][3], BF16_TOL, "tc LN stats vs simt")
