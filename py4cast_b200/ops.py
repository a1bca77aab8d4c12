"""
torch.autograd wrappers around the C-ABI kernels for the graph-model path:
CSR construction, fused MLP (+LayerNorm/residual) and the InteractionNet
message-passing step.  PyTorch is used for memory, streams and autograd
bookkeeping only; every FLOP of the forward/backward runs in libpy4cast_b200.

Reference semantics replaced: mfai `make_mlp` / `InteractionNet` (SURVEY.md §3.2,
Appendix A.1-A.2), reached from py4cast/lightning.py:593-596 (`self.model(x)`).
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import HIDDEN, ArX, Mlp, MlpBwd, check, dtype_code, ptr, stream

# ------------------------------------------------------------------ workspaces
_WGRAD_WS: dict = {}
_WGRAD_WS_FLOATS = 2 * 148 * (192 * 64 + 64 * 64 + 4 * 64)


def _wgrad_workspace(device: torch.device, floats: int) -> torch.Tensor:
    ws = _WGRAD_WS.get(device)
    need = max(floats, _WGRAD_WS_FLOATS)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.float32, device=device)
        _WGRAD_WS[device] = ws
    return ws


def _esize(t: torch.Tensor) -> int:
    return t.element_size()


def _off(t: torch.Tensor, elements: int) -> int:
    """device pointer of `t` advanced by `elements` elements"""
    return ptr(t) + elements * _esize(t)


# -------------------------------------------------------------------------- CSR
@dataclass
class Csr:
    rowptr: torch.Tensor  # int32 [num_seg+1]
    perm: Optional[torch.Tensor]  # int32 [E] or None when segments are contiguous (identity)
    num_seg: int
    inv_deg: Optional[torch.Tensor] = None  # float32 [num_seg] = 1/max(1,deg)


def csr_build(keys: torch.Tensor, num_keys: int, drop_identity_perm: bool = True) -> Csr:
    """Group edge ids by key (stable): bit-exact with torch.sort(keys, stable=True)."""
    if keys.dtype != torch.int64:
        raise TypeError("csr_build expects int64 keys")
    keys = keys.contiguous()
    E = keys.numel()
    dev = keys.device
    rowptr = torch.empty(num_keys + 1, dtype=torch.int32, device=dev)
    perm = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
    scratch = torch.empty(E + num_keys + 1, dtype=torch.int32, device=dev)
    check(
        _lib.load().p4c_csr_build(
            ptr(keys), E, num_keys, ptr(rowptr), ptr(perm), ptr(scratch), stream()
        ),
        "p4c_csr_build",
    )
    perm = perm[:E]
    deg = (rowptr[1:] - rowptr[:-1]).clamp(min=1).to(torch.float32)
    identity = False
    if drop_identity_perm and E > 0:
        identity = bool(torch.equal(perm, torch.arange(E, dtype=torch.int32, device=dev)))
    return Csr(rowptr, None if identity else perm, num_keys, (1.0 / deg).contiguous())


def index_to_i32(idx: torch.Tensor) -> torch.Tensor:
    idx = idx.contiguous()
    out = torch.empty(idx.numel(), dtype=torch.int32, device=idx.device)
    if idx.numel():
        check(_lib.load().p4c_index_to_i32(ptr(idx), ptr(out), idx.numel(), stream()), "p4c_index_to_i32")
    return out


@dataclass
class EdgeGroup:
    """A contiguous range of edges sharing one set of edge-MLP weights."""

    start: int
    stop: int
    by_rec: Csr  # edges of this group keyed by receiver (global edge ids)
    by_send: Csr


@dataclass
class GraphIndex:
    """Device-side index structures of one edge set (built once at model set-up)."""

    num_send: int
    num_rec: int
    num_edges: int
    send_idx: torch.Tensor  # int32 [E]
    rec_idx: torch.Tensor  # int32 [E]
    by_rec: Csr  # all edges keyed by receiver
    groups: List[EdgeGroup] = field(default_factory=list)
    node_chunks: Optional[List[Tuple[int, int]]] = None  # receiver ranges with own node-MLP weights

    @staticmethod
    def build(
        edge_index: torch.Tensor,
        num_send: int,
        num_rec: int,
        edge_chunk_sizes: Optional[Sequence[int]] = None,
        node_chunk_sizes: Optional[Sequence[int]] = None,
    ) -> "GraphIndex":
        assert edge_index.is_cuda and edge_index.dtype == torch.int64
        send, rec = edge_index[0].contiguous(), edge_index[1].contiguous()
        E = int(send.numel())
        by_rec = csr_build(rec, num_rec)
        groups = []
        if edge_chunk_sizes is None:
            groups.append(EdgeGroup(0, E, by_rec, csr_build(send, num_send)))
        else:
            start = 0
            for n in edge_chunk_sizes:
                stop = start + int(n)
                # keys outside the chunk are parked on an extra trailing segment
                def keyed(keys, nseg):
                    k = torch.full_like(keys, nseg)
                    k[start:stop] = keys[start:stop]
                    c = csr_build(k, nseg + 1, drop_identity_perm=False)
                    return Csr(c.rowptr[: nseg + 1].contiguous(), c.perm, nseg, c.inv_deg[:nseg].contiguous())

                groups.append(EdgeGroup(start, stop, keyed(rec, num_rec), keyed(send, num_send)))
                start = stop
            assert start == E, "edge_chunk_sizes must cover all edges"
        node_chunks = None
        if node_chunk_sizes is not None:
            node_chunks, s = [], 0
            for n in node_chunk_sizes:
                node_chunks.append((s, s + int(n)))
                s += int(n)
            assert s == num_rec
        return GraphIndex(
            num_send, num_rec, E, index_to_i32(send), index_to_i32(rec), by_rec, groups, node_chunks
        )


# ----------------------------------------------------------------- kernel calls
def segment_sum(src: torch.Tensor, csr: Csr, batch: int, mean: bool = False,
                out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    """src (B,E,64) -> out (B,num_seg,64)."""
    assert src.dim() == 3 and src.shape[-1] == HIDDEN and src.is_contiguous()
    if out is None:
        out = torch.empty(batch, csr.num_seg, HIDDEN, dtype=src.dtype, device=src.device)
    check(
        _lib.load().p4c_segment_sum(
            ptr(src), src.shape[1] * HIDDEN, ptr(csr.rowptr), ptr(csr.perm), csr.num_seg, batch,
            ptr(out), out.shape[1] * HIDDEN, int(mean), int(accumulate), dtype_code(src.dtype), stream(),
        ),
        "p4c_segment_sum",
    )
    return out


def rows_matmul(x: torch.Tensor, w1: torch.Tensor, col0: int, out: torch.Tensor, accumulate: bool) -> None:
    """out[b,r,:] (+)= x[b,r,:] @ w1[:, col0:col0+64]  (w1 is the nn.Linear weight [64, k1])."""
    B, R, _ = x.shape
    check(
        _lib.load().p4c_rows_matmul(
            ptr(x), R * HIDDEN, ptr(w1), w1.shape[1], col0, R, B, ptr(out), out.shape[1] * HIDDEN,
            int(accumulate), dtype_code(x.dtype), stream(),
        ),
        "p4c_rows_matmul",
    )


def batch_sum(x: torch.Tensor) -> torch.Tensor:
    """(B, ...) -> (...) summed over the batch, fixed order."""
    if x.shape[0] == 1:
        return x[0]
    out = torch.empty(x.shape[1:], dtype=x.dtype, device=x.device)
    n = out.numel()
    check(
        _lib.load().p4c_batch_sum(ptr(x), n, x.shape[0], n, ptr(out), 0, dtype_code(x.dtype), stream()),
        "p4c_batch_sum",
    )
    return out


def cast(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    if x.dtype == dtype:
        return x
    x = x.contiguous()
    out = torch.empty(x.shape, dtype=dtype, device=x.device)
    check(
        _lib.load().p4c_cast(ptr(x), dtype_code(x.dtype), ptr(out), dtype_code(dtype), x.numel(), stream()),
        "p4c_cast",
    )
    return out


class _CastFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dtype):
        ctx.src_dtype = x.dtype
        return cast(x, dtype)

    @staticmethod
    def backward(ctx, g):
        return cast(g.contiguous(), ctx.src_dtype), None


def cast_autograd(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    return x if x.dtype == dtype else _CastFn.apply(x, dtype)


# ------------------------------------------------------------------- descriptors
@dataclass
class _SrcSpec:
    tensor: torch.Tensor  # (B,R,W) or (R,W) shared across the batch
    index: Optional[torch.Tensor] = None  # int32 gather index (already offset to the row range)
    row_offset: int = 0  # first row (identity sources covering a sub-range)
    index_offset: int = 0  # first entry of `index` used


def _fill_mlp(desc: Mlp, rows: int, batch: int, full_rows: int, srcs: Sequence[_SrcSpec],
              weights: Sequence[Optional[torch.Tensor]], out_features: int, dtype: torch.dtype) -> None:
    w1, b1, w2, b2, gamma, beta = weights
    desc.rows, desc.batch, desc.nsrc = rows, batch, len(srcs)
    desc.out_features = out_features
    desc.layer_norm = int(gamma is not None)
    desc.dtype = dtype_code(dtype)
    desc.full_rows = full_rows
    k1 = 0
    for i, s in enumerate(srcs):
        t = s.tensor
        assert t.dtype == dtype, f"source {i} dtype {t.dtype} != {dtype}"
        assert t.is_contiguous()
        width = t.shape[-1]
        shared = t.dim() == 2
        n_rows = t.shape[-2]
        desc.src[i].ptr = _off(t, s.row_offset * width)
        desc.src[i].index = None if s.index is None else _off(s.index, s.index_offset)
        desc.src[i].batch_stride = 0 if shared else n_rows * width
        desc.src[i].ld = width
        desc.src[i].width = width
        k1 += width
    desc.k1 = k1
    assert w1.shape == (HIDDEN, k1), f"w1 shape {tuple(w1.shape)} != (64,{k1})"
    assert w2.shape == (out_features, HIDDEN)
    desc.w1, desc.b1, desc.w2, desc.b2 = ptr(w1), ptr(b1), ptr(w2), ptr(b2)
    desc.gamma, desc.beta = ptr(gamma), ptr(beta)
    desc.ln_eps = 1e-5


def _grad_target(p: Optional[torch.Tensor], needed: bool):
    """(buffer to accumulate into, value to return from backward)."""
    if p is None or not needed:
        return None, None
    sink = getattr(p, "_p4c_grad", None)
    if sink is not None:
        return sink, None
    buf = torch.zeros_like(p, dtype=torch.float32)
    return buf, buf


def _check_hidden(t: torch.Tensor, what: str) -> None:
    if t.shape[-1] != HIDDEN:
        raise NotImplementedError(
            f"{what}: the fused kernels are built for hidden_dims={HIDDEN}, got {t.shape[-1]}"
        )


# -------------------------------------------------------------------- fused MLP
class _MLPFn(torch.autograd.Function):
    """y = [LN](W2 silu(W1 x + b1) + b2) [+ x]   for one dense, identity-indexed source."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, gamma, beta, residual: bool):
        lib = _lib.load()
        shared = x.dim() == 2
        x = x.contiguous()
        B = 1 if shared else x.shape[0]
        R, K = x.shape[-2], x.shape[-1]
        out_features = w2.shape[0]
        if w1.shape[0] != HIDDEN:
            raise NotImplementedError(f"fused MLP needs hidden width {HIDDEN}, got {w1.shape[0]}")
        if residual and (K != HIDDEN or out_features != HIDDEN):
            raise ValueError("residual needs 64-wide input and output")
        dt = x.dtype
        need_grad = any(ctx.needs_input_grad)
        out = torch.empty((R, out_features) if shared else (B, R, out_features), dtype=dt, device=x.device)
        h1 = torch.empty(B * R, HIDDEN, dtype=dt, device=x.device) if need_grad else None
        stats = (
            torch.empty(B * R, 2, dtype=torch.float32, device=x.device)
            if (need_grad and gamma is not None)
            else None
        )
        desc = Mlp()
        _fill_mlp(desc, R, B, 0, [_SrcSpec(x)], (w1, b1, w2, b2, gamma, beta), out_features, dt)
        check(
            lib.p4c_mlp_forward(C.byref(desc), ptr(x) if residual else None, ptr(out), None,
                                ptr(h1), ptr(stats), stream()),
            "p4c_mlp_forward",
        )
        ctx.save_for_backward(x, w1, b1, w2, b2, gamma, beta, h1, stats)
        ctx.residual = residual
        return out

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, w1, b1, w2, b2, gamma, beta, h1, stats = ctx.saved_tensors
        shared = x.dim() == 2
        B = 1 if shared else x.shape[0]
        R, K = x.shape[-2], x.shape[-1]
        out_features = w2.shape[0]
        dt = x.dtype
        dy = dy.contiguous()
        if dy.dtype != dt:
            dy = cast(dy, dt)
        desc = Mlp()
        _fill_mlp(desc, R, B, 0, [_SrcSpec(x)], (w1, b1, w2, b2, gamma, beta), out_features, dt)
        g = MlpBwd()
        g.dy_a, g.dy_a_batch_stride = ptr(dy), R * out_features
        g.h1, g.stats = ptr(h1), ptr(stats)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            g.dx[0], g.dx_ld[0], g.dx_accumulate[0] = ptr(dx), K, 0
            g.dx0_add_dy = 2 if ctx.residual else 0
        elif ctx.residual:
            raise RuntimeError("residual MLP with a non-differentiable input is not supported")
        rets = []
        for p, slot, needed in (
            (w1, "g_w1", ctx.needs_input_grad[1]), (b1, "g_b1", ctx.needs_input_grad[2]),
            (w2, "g_w2", ctx.needs_input_grad[3]), (b2, "g_b2", ctx.needs_input_grad[4]),
            (gamma, "g_gamma", ctx.needs_input_grad[5]), (beta, "g_beta", ctx.needs_input_grad[6]),
        ):
            buf, ret = _grad_target(p, needed)
            setattr(g, slot, ptr(buf))
            rets.append(ret)
        nfl = lib.p4c_mlp_wgrad_partial_floats(C.byref(desc))
        ws = _wgrad_workspace(x.device, nfl)
        g.wgrad_partials, g.wgrad_partials_floats = ptr(ws), ws.numel()
        check(lib.p4c_mlp_backward(C.byref(desc), C.byref(g), stream()), "p4c_mlp_backward")
        return (dx, *rets, None)


def fused_mlp(x: torch.Tensor, weights: Sequence[Optional[torch.Tensor]], residual: bool = False) -> torch.Tensor:
    w1, b1, w2, b2, gamma, beta = weights
    return _MLPFn.apply(x, w1, b1, w2, b2, gamma, beta, residual)


# --------------------------------------------------------------- InteractionNet
def _rows(t: torch.Tensor) -> int:
    return t.shape[-2]


class _InteractionNetFn(torch.autograd.Function):
    """
    (send, rec, edge) -> (rec_out, edge_out):
        msg      = edge_mlp(cat(edge, send[sender], rec[receiver]))
        aggr     = segment_{sum|mean}(msg by receiver)
        rec_out  = rec + aggr_mlp(cat(rec, aggr))
        edge_out = edge + msg                       (update_edges)
    Weights: 6 tensors per edge group followed by 6 per node chunk.
    """

    @staticmethod
    def forward(ctx, gi: GraphIndex, update_edges: bool, mean: bool, send, rec, edge, *weights):
        lib = _lib.load()
        for t, n in ((send, "send"), (rec, "rec"), (edge, "edge")):
            _check_hidden(t, f"InteractionNet {n} representation")
        send, rec, edge = send.contiguous(), rec.contiguous(), edge.contiguous()
        dt = rec.dtype
        assert send.dtype == dt and edge.dtype == dt
        B = max([t.shape[0] for t in (send, rec, edge) if t.dim() == 3] or [1])
        E, Nr, Ns = gi.num_edges, gi.num_rec, gi.num_send
        assert _rows(edge) == E and _rows(rec) == Nr and _rows(send) == Ns, (
            f"InteractionNet shapes: edge {tuple(edge.shape)} rec {tuple(rec.shape)} send "
            f"{tuple(send.shape)} vs graph E={E} Nr={Nr} Ns={Ns}"
        )
        n_eg = len(gi.groups)
        node_chunks = gi.node_chunks or [(0, Nr)]
        ew = [weights[6 * i: 6 * i + 6] for i in range(n_eg)]
        nw = [weights[6 * (n_eg + i): 6 * (n_eg + i) + 6] for i in range(len(node_chunks))]
        dev = rec.device
        need_grad = any(ctx.needs_input_grad)
        ctx.set_materialize_grads(False)

        msg = torch.empty(B, E, HIDDEN, dtype=dt, device=dev)
        edge_out = torch.empty(B, E, HIDDEN, dtype=dt, device=dev) if update_edges else None
        h1_e = torch.empty(B, E, HIDDEN, dtype=dt, device=dev) if need_grad else None
        st_e = torch.empty(B, E, 2, dtype=torch.float32, device=dev) if need_grad else None
        for grp, w in zip(gi.groups, ew):
            n = grp.stop - grp.start
            if n == 0:
                continue
            desc = Mlp()
            _fill_mlp(
                desc, n, B, E,
                [
                    _SrcSpec(edge, row_offset=grp.start),
                    _SrcSpec(send, index=gi.send_idx, index_offset=grp.start),
                    _SrcSpec(rec, index=gi.rec_idx, index_offset=grp.start),
                ],
                w, HIDDEN, dt,
            )
            o = grp.start * HIDDEN
            check(
                lib.p4c_mlp_forward(
                    C.byref(desc), None, _off(msg, o),
                    _off(edge_out, o) if update_edges else None,
                    _off(h1_e, o) if need_grad else None,
                    _off(st_e, grp.start * 2) if need_grad else None, stream(),
                ),
                "p4c_mlp_forward(edge)",
            )
        aggr = segment_sum(msg, gi.by_rec, B, mean=mean)
        del msg

        rec_b = rec if rec.dim() == 3 else rec.unsqueeze(0).expand(B, Nr, HIDDEN).contiguous()
        rec_out = torch.empty(B, Nr, HIDDEN, dtype=dt, device=dev)
        h1_n = torch.empty(B, Nr, HIDDEN, dtype=dt, device=dev) if need_grad else None
        st_n = torch.empty(B, Nr, 2, dtype=torch.float32, device=dev) if need_grad else None
        for (n0, n1), w in zip(node_chunks, nw):
            desc = Mlp()
            _fill_mlp(
                desc, n1 - n0, B, Nr,
                [_SrcSpec(rec_b, row_offset=n0), _SrcSpec(aggr, row_offset=n0)],
                w, HIDDEN, dt,
            )
            o = n0 * HIDDEN
            check(
                lib.p4c_mlp_forward(
                    C.byref(desc), _off(rec_b, o), _off(rec_out, o), None,
                    _off(h1_n, o) if need_grad else None,
                    _off(st_n, n0 * 2) if need_grad else None, stream(),
                ),
                "p4c_mlp_forward(node)",
            )
        ctx.gi, ctx.update_edges, ctx.mean, ctx.B = gi, update_edges, mean, B
        ctx.rec_shared = rec.dim() == 2
        ctx.n_weights = len(weights)
        ctx.save_for_backward(send, rec_b, edge, aggr, h1_e, st_e, h1_n, st_n, *weights)
        if update_edges:
            return rec_out, edge_out
        return rec_out

    @staticmethod
    def backward(ctx, d_rec_out, d_edge_out=None):
        lib = _lib.load()
        gi: GraphIndex = ctx.gi
        saved = ctx.saved_tensors
        send, rec_b, edge, aggr, h1_e, st_e, h1_n, st_n = saved[:8]
        weights = saved[8:]
        B, E, Nr, Ns = ctx.B, gi.num_edges, gi.num_rec, gi.num_send
        dt, dev = rec_b.dtype, rec_b.device
        n_eg = len(gi.groups)
        node_chunks = gi.node_chunks or [(0, Nr)]
        ew = [weights[6 * i: 6 * i + 6] for i in range(n_eg)]
        nw = [weights[6 * (n_eg + i): 6 * (n_eg + i) + 6] for i in range(len(node_chunks))]
        need_send, need_rec, need_edge = ctx.needs_input_grad[3:6]
        wneed = ctx.needs_input_grad[6:]

        if d_rec_out is None:
            d_rec_out = torch.zeros(B, Nr, HIDDEN, dtype=dt, device=dev)
        d_rec_out = d_rec_out.contiguous()
        if d_rec_out.dtype != dt:
            d_rec_out = cast(d_rec_out, dt)
        if d_edge_out is not None:
            d_edge_out = d_edge_out.contiguous()
            if d_edge_out.dtype != dt:
                d_edge_out = cast(d_edge_out, dt)

        wgrads: List[Optional[torch.Tensor]] = [None] * len(weights)

        def bind_wgrads(g: MlpBwd, base: int, ws: Sequence[Optional[torch.Tensor]]):
            for j, slot in enumerate(("g_w1", "g_b1", "g_w2", "g_b2", "g_gamma", "g_beta")):
                buf, ret = _grad_target(ws[j], wneed[base + j])
                setattr(g, slot, ptr(buf))
                wgrads[base + j] = ret

        # ---- node MLP backward: d_rec (incl. residual pass-through) and d_aggr
        d_rec = torch.empty(B, Nr, HIDDEN, dtype=dt, device=dev)
        d_aggr = torch.empty(B, Nr, HIDDEN, dtype=dt, device=dev)
        for ci, ((n0, n1), w) in enumerate(zip(node_chunks, nw)):
            desc = Mlp()
            _fill_mlp(
                desc, n1 - n0, B, Nr,
                [_SrcSpec(rec_b, row_offset=n0), _SrcSpec(aggr, row_offset=n0)], w, HIDDEN, dt,
            )
            o = n0 * HIDDEN
            g = MlpBwd()
            g.dy_a, g.dy_a_batch_stride = _off(d_rec_out, o), Nr * HIDDEN
            g.h1, g.stats = _off(h1_n, o), _off(st_n, n0 * 2)
            g.dx[0], g.dx_ld[0], g.dx_accumulate[0] = _off(d_rec, o), HIDDEN, 0
            g.dx[1], g.dx_ld[1], g.dx_accumulate[1] = _off(d_aggr, o), HIDDEN, 0
            g.dx0_add_dy = 2
            bind_wgrads(g, 6 * (n_eg + ci), w)
            nfl = lib.p4c_mlp_wgrad_partial_floats(C.byref(desc))
            ws = _wgrad_workspace(dev, nfl)
            g.wgrad_partials, g.wgrad_partials_floats = ptr(ws), ws.numel()
            check(lib.p4c_mlp_backward(C.byref(desc), C.byref(g), stream()), "p4c_mlp_backward(node)")

        # ---- edge MLP backward
        dh1 = torch.empty(B, E, HIDDEN, dtype=dt, device=dev)
        d_edge = torch.empty(B, E, HIDDEN, dtype=dt, device=dev) if need_edge else None
        d_send = torch.empty(B, Ns, HIDDEN, dtype=dt, device=dev) if need_send else None
        first_send = True
        for gidx, (grp, w) in enumerate(zip(gi.groups, ew)):
            n = grp.stop - grp.start
            if n == 0:
                continue
            desc = Mlp()
            _fill_mlp(
                desc, n, B, E,
                [
                    _SrcSpec(edge, row_offset=grp.start),
                    _SrcSpec(send, index=gi.send_idx, index_offset=grp.start),
                    _SrcSpec(rec_b, index=gi.rec_idx, index_offset=grp.start),
                ],
                w, HIDDEN, dt,
            )
            o = grp.start * HIDDEN
            g = MlpBwd()
            g.dy_a = ptr(d_aggr)
            g.dy_a_index = _off(gi.rec_idx, grp.start)
            g.dy_a_scale = ptr(gi.by_rec.inv_deg) if ctx.mean else None
            g.dy_a_batch_stride = Nr * HIDDEN
            g.dy_b = _off(d_edge_out, o) if d_edge_out is not None else None
            g.h1, g.stats = _off(h1_e, o), _off(st_e, grp.start * 2)
            g.dh1_out = _off(dh1, o)
            if need_edge:
                g.dx[0], g.dx_ld[0], g.dx_accumulate[0] = _off(d_edge, o), HIDDEN, 0
                g.dx0_add_dy = 1 if d_edge_out is not None else 0
            bind_wgrads(g, 6 * gidx, w)
            nfl = lib.p4c_mlp_wgrad_partial_floats(C.byref(desc))
            ws = _wgrad_workspace(dev, nfl)
            g.wgrad_partials, g.wgrad_partials_floats = ptr(ws), ws.numel()
            check(lib.p4c_mlp_backward(C.byref(desc), C.byref(g), stream()), "p4c_mlp_backward(edge)")
            # sender / receiver gradients by linearity: (segment-sum of dh1) x W1 slice
            if need_rec:
                s_r = segment_sum(dh1, grp.by_rec, B)
                rows_matmul(s_r, w[0], 2 * HIDDEN, d_rec, accumulate=True)
            if need_send:
                s_s = segment_sum(dh1, grp.by_send, B)
                rows_matmul(s_s, w[0], HIDDEN, d_send, accumulate=not first_send)
                first_send = False
        if need_send and first_send:
            d_send.zero_()

        g_send = (batch_sum(d_send) if send.dim() == 2 else d_send) if need_send else None
        g_rec = (batch_sum(d_rec) if ctx.rec_shared else d_rec) if need_rec else None
        g_edge = (batch_sum(d_edge) if edge.dim() == 2 else d_edge) if need_edge else None
        return (None, None, None, g_send, g_rec, g_edge, *wgrads)


def interaction_net(gi: GraphIndex, send, rec, edge, weights: Sequence[torch.Tensor],
                    update_edges: bool, mean: bool):
    return _InteractionNetFn.apply(gi, update_edges, mean, send, rec, edge, *weights)
