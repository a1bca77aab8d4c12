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
from ._lib import HIDDEN, PACK_BYTES, ArX, Mlp, MlpBwd, Pack, SegPart, WGrad, check, dtype_code, ptr, stream

# ------------------------------------------------------------------ workspaces
_WGRAD_WS: dict = {}
_WGRAD_WS_FLOATS = 2 * 148 * (192 * 64 + 64 * 64 + 4 * 64)


def _wgrad_workspace(key, floats: int) -> torch.Tensor:
    """Shared scratch for in-call weight-gradient reductions, per (device, execution branch)."""
    ws = _WGRAD_WS.get(key)
    need = max(floats, _WGRAD_WS_FLOATS)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.float32, device=key[0])
        _WGRAD_WS[key] = ws
    return ws


def _esize(t: torch.Tensor) -> int:
    return t.element_size()


def _off(t: torch.Tensor, elements: int) -> int:
    """device pointer of `t` advanced by `elements` elements"""
    return ptr(t) + elements * _esize(t)


# --------------------------------------------------- per-MLP side state (tensor-core engine)
def _table_to_device(entries, device: torch.device) -> torch.Tensor:
    """ctypes structs -> one uint8 device tensor (a DEVICE table for the batched kernels)."""
    raw = b"".join(bytes(e) for e in entries)
    return torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)


class MlpMeta:
    """Side state of one 2-layer MLP: the packed bf16 weight image (p4c_pack_weights) and the
    per-CTA weight-gradient partials whose reduction is deferred to ONE batched launch per
    backward pass (`flush_wgrads`).  Both are optional accelerations: a call without meta converts
    the weights in the kernel and reduces its partials itself."""

    def __init__(self, params: Sequence[Optional[torch.Tensor]], widths: Sequence[int], out_features: int):
        self.params = tuple(params)  # w1, b1, w2, b2, gamma, beta
        self.widths = tuple(int(w) for w in widths)
        self.k1 = int(sum(self.widths))
        self.out_features = int(out_features)
        self.image: Optional[torch.Tensor] = None  # uint8 [PACK_BYTES] view, valid after pack()
        # deferred weight-gradient partials, one set per execution branch (stream): branch -> tensor
        self.partials: dict = {}
        self.partials_blocks: dict = {}
        self.partials_used: dict = {}

    def image_ptr(self, dtype: torch.dtype) -> Optional[int]:
        if self.image is None or dtype != torch.bfloat16:
            return None
        return self.image.data_ptr()

    def sinks(self) -> Optional[List[Optional[torch.Tensor]]]:
        """The flat-gradient views of the parameters, if every parameter has one."""
        out = []
        for p in self.params:
            if p is None:
                out.append(None)
                continue
            g = getattr(p, "_p4c_grad", None)
            if g is None or not p.requires_grad:
                return None
            out.append(g)
        return out


class WeightPacker:
    """Packs the weights of a set of MLPs into their images with ONE kernel launch."""

    def __init__(self, metas: Sequence[MlpMeta]):
        self.metas = list(metas)
        self._images: Optional[torch.Tensor] = None
        self._table: Optional[torch.Tensor] = None
        self._key = None

    def _build(self, device: torch.device) -> None:
        n = len(self.metas)
        self._images = torch.empty(n, PACK_BYTES, dtype=torch.uint8, device=device)
        entries = []
        for i, m in enumerate(self.metas):
            e = Pack()
            w1, b1, w2, b2, gamma, beta = m.params
            e.w1, e.b1, e.w2, e.b2 = ptr(w1), ptr(b1), ptr(w2), ptr(b2)
            e.gamma, e.beta = ptr(gamma), ptr(beta)
            e.dst = self._images[i].data_ptr()
            e.k1, e.out_features, e.nsrc = m.k1, m.out_features, len(m.widths)
            for j, w in enumerate(m.widths):
                e.width[j] = w
            entries.append(e)
            m.image = self._images[i]
        self._table = _table_to_device(entries, device)

    def pack(self) -> None:
        if not self.metas:
            return
        w = self.metas[0].params[0]
        key = tuple(p.data_ptr() for m in self.metas for p in m.params if p is not None)
        if self._table is None or key != self._key:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("WeightPacker: parameters moved; run one eager step before CUDA-graph capture")
            self._build(w.device)
            self._key = key
        check(_lib.load().p4c_pack_weights(ptr(self._table), len(self.metas), stream()), "p4c_pack_weights")


_DEFERRED: List[Tuple[MlpMeta, int]] = []  # (meta, branch) touched since the last flush
_DEFERRED_TABLES: dict = {}
_BRANCH = 0


def set_branch(i: int) -> None:
    """Execution branch (one per concurrently running micro-batch / CUDA stream) of the autograd
    nodes created from now on: each branch accumulates weight-gradient partials in its own buffers."""
    global _BRANCH
    _BRANCH = int(i)


def current_branch() -> int:
    return _BRANCH


def _deferred_mode(meta: Optional[MlpMeta], nblocks: int, partial_floats: int, device, branch: int = 0) -> int:
    """0 = reduce in the call; 1 / 2 = overwrite / accumulate this MLP's own partials (deferred)."""
    if meta is None or meta.sinks() is None:
        return 0
    if branch not in meta.partials:
        if torch.cuda.is_current_stream_capturing():
            return 0
        meta.partials[branch] = torch.empty(partial_floats, dtype=torch.float32, device=device)
        meta.partials_blocks[branch] = nblocks
    if meta.partials_blocks[branch] != nblocks or meta.partials[branch].numel() < partial_floats:
        return 0  # same weights used with another problem size: keep that call self-contained
    if not meta.partials_used.get(branch, False):
        meta.partials_used[branch] = True
        _DEFERRED.append((meta, branch))
        return 1
    return 2


def reset_deferred_wgrads() -> None:
    """Drops deferred weight-gradient state left behind by a backward pass that was never flushed
    (a `loss.backward()` outside the trainer, or a step that raised before `flush_wgrads`): the
    next backward then OVERWRITES its partials instead of accumulating onto stale ones.  Host-side
    flags only (safe before or under CUDA-graph capture); the trainer calls it at the start of
    every step."""
    for meta, branch in _DEFERRED:
        meta.partials_used[branch] = False
    _DEFERRED.clear()


def flush_wgrads() -> None:
    """Reduces the deferred weight-gradient partials of every MLP touched since the last flush
    into the parameters' gradient sinks: ONE launch per execution branch, branches in ascending
    order (fixed summation order), issued on the current stream (the trainer calls it after
    backward, when every branch has been joined)."""
    if not _DEFERRED:
        return
    for branch in sorted({b for _, b in _DEFERRED}):
        metas = [m for m, b in _DEFERRED if b == branch]
        sig = tuple((m.partials[branch].data_ptr(),) + tuple(g.data_ptr() for g in m.sinks() if g is not None)
                    for m in metas)
        key = (branch, tuple(id(m) for m in metas))
        ent = _DEFERRED_TABLES.get(key)
        if ent is None or ent[1] != sig:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("flush_wgrads: run one eager step before CUDA-graph capture")
            entries = []
            for m in metas:
                e = WGrad()
                e.partials = ptr(m.partials[branch])
                for name, g in zip(("g_w1", "g_b1", "g_w2", "g_b2", "g_gamma", "g_beta"), m.sinks()):
                    setattr(e, name, ptr(g))
                e.nblocks, e.k1, e.out_features = m.partials_blocks[branch], m.k1, m.out_features
                entries.append(e)
            ent = (_table_to_device(entries, metas[0].partials[branch].device), sig, metas)
            _DEFERRED_TABLES[key] = ent
        check(_lib.load().p4c_wgrad_reduce_batched(ptr(ent[0]), len(metas), stream()), "p4c_wgrad_reduce_batched")
        for m in metas:
            m.partials_used[branch] = False
    _DEFERRED.clear()


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
    # > 0: edges are sorted by receiver and every receiver has exactly this many (1, 2, 4 or 8):
    # edge e belongs to receiver e // uniform_deg, so the aggregation can run in the
    # epilogue of the edge kernel (m2g: 4 nearest mesh nodes; up / down edges: 1)
    uniform_deg: int = 0

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
                # CSR of the chunk's own edges; perm holds GLOBAL edge ids (local + start)
                def keyed(keys, nseg):
                    c = csr_build(keys[start:stop].contiguous(), nseg, drop_identity_perm=False)
                    return Csr(c.rowptr, (c.perm + start).contiguous(), nseg, c.inv_deg)

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
        uniform = 0
        if edge_chunk_sizes is None and by_rec.perm is None and num_rec > 0 and E % num_rec == 0:
            d = E // num_rec
            if d in (1, 2, 4, 8):
                expect = torch.arange(num_rec + 1, dtype=torch.int32, device=rec.device) * d
                if bool(torch.equal(by_rec.rowptr, expect)):
                    uniform = d
        return GraphIndex(
            num_send, num_rec, E, index_to_i32(send), index_to_i32(rec), by_rec, groups, node_chunks, uniform
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
    tensor: torch.Tensor  # (B,R,W) or (R,W) shared across the batch; rows may be strided (ld >= W)
    index: Optional[torch.Tensor] = None  # int32 gather index (already offset to the row range)
    row_offset: int = 0  # first row (identity sources covering a sub-range)
    index_offset: int = 0  # first entry of `index` used
    seg: Optional[Csr] = None  # segment-sum source: input row r = sum of tensor rows of segment r
    seg_row_offset: int = 0  # first segment used
    seg_mean: bool = False
    seg_out: Optional[torch.Tensor] = None  # (B, full_rows, 64) receives the summed rows


def _row_layout(t: torch.Tensor) -> Tuple[int, int]:
    """(ld, batch_stride) in elements of a (B,R,W) / (R,W) tensor with unit-stride features."""
    assert t.stride(-1) == 1, "feature dimension must be contiguous"
    ld = t.stride(-2) if t.shape[-2] > 1 else t.shape[-1]
    if t.dim() == 2:
        return ld, 0
    bs = t.stride(0) if t.shape[0] > 1 else t.shape[-2] * ld
    return ld, bs


def _rows_ok(t: torch.Tensor) -> bool:
    """Row-strided layouts the kernels take without a copy: unit feature stride, ld >= width."""
    if t.dim() not in (2, 3) or t.stride(-1) != 1:
        return False
    if t.shape[-2] > 1 and t.stride(-2) < t.shape[-1]:
        return False
    return True


def _as_rows(t: torch.Tensor) -> torch.Tensor:
    return t if _rows_ok(t) else t.contiguous()


def _fill_mlp(desc: Mlp, rows: int, batch: int, full_rows: int, srcs: Sequence[_SrcSpec],
              weights: Sequence[Optional[torch.Tensor]], out_features: int, dtype: torch.dtype,
              meta: Optional[MlpMeta] = None) -> None:
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
        width = t.shape[-1]
        ld, bs = _row_layout(t)
        desc.src[i].ptr = _off(t, s.row_offset * ld)
        desc.src[i].index = None if s.index is None else _off(s.index, s.index_offset)
        desc.src[i].batch_stride = bs
        desc.src[i].ld = ld
        desc.src[i].width = width
        if s.seg is not None:
            assert width == HIDDEN and ld == HIDDEN and s.index is None
            desc.src[i].seg_rowptr = _off(s.seg.rowptr, s.seg_row_offset)
            desc.src[i].seg_perm = ptr(s.seg.perm)
            desc.src[i].seg_scale = _off(s.seg.inv_deg, s.seg_row_offset) if s.seg_mean else None
            desc.src[i].seg_out = None if s.seg_out is None else _off(s.seg_out, s.seg_row_offset * HIDDEN)
        else:
            desc.src[i].seg_rowptr = desc.src[i].seg_perm = desc.src[i].seg_scale = desc.src[i].seg_out = None
        k1 += width
    desc.k1 = k1
    assert w1.shape == (HIDDEN, k1), f"w1 shape {tuple(w1.shape)} != (64,{k1})"
    assert w2.shape == (out_features, HIDDEN)
    desc.w1, desc.b1, desc.w2, desc.b2 = ptr(w1), ptr(b1), ptr(w2), ptr(b2)
    desc.gamma, desc.beta = ptr(gamma), ptr(beta)
    desc.ln_eps = 1e-5
    desc.packed = None
    desc.tile_rows = 0
    for sp in srcs:
        if sp.seg is not None:
            desc.tile_rows = _seg_tile_rows(sp.tensor.shape[-2], sp.seg.num_seg)
    if meta is not None:
        assert meta.widths == tuple(s.tensor.shape[-1] for s in srcs), "weight image packed for other source widths"
        desc.packed = meta.image_ptr(dtype)


def _seg_tile_rows(num_edges: int, num_seg: int) -> int:
    """Valid rows per CTA tile of a segment-sum prologue: ~1k edges per CTA, so that long segments
    (g2m: ~70 edges per mesh node, m2g by sender: ~200) are spread over many CTAs."""
    avg = max(1.0, num_edges / max(1, num_seg))
    rows = 128
    while rows > 16 and rows * avg > 1024:
        rows //= 2
    return rows


def _grad_target(p: Optional[torch.Tensor], needed: bool):
    """(buffer to accumulate into, value to return from backward)."""
    if p is None or not needed:
        return None, None
    sink = getattr(p, "_p4c_grad", None)
    if sink is not None:
        return sink, None
    buf = torch.zeros_like(p, dtype=torch.float32)
    return buf, buf


def _bind_wgrads(lib, desc: Mlp, g: MlpBwd, ws: Sequence[Optional[torch.Tensor]], needed: Sequence[bool],
                 meta: Optional[MlpMeta], device, branch: int = 0) -> List[Optional[torch.Tensor]]:
    """Weight-gradient outputs of one backward call: deferred partials when the MLP has gradient
    sinks (trainer), otherwise the shared workspace + in-call reduction.  Returns the 6 values to
    hand back to autograd."""
    nfl = lib.p4c_mlp_wgrad_partial_floats(C.byref(desc))
    rets: List[Optional[torch.Tensor]] = [None] * 6
    mode = 0
    if desc.dtype == _lib.P4C_BF16 and all(needed[j] or ws[j] is None for j in range(6)):
        mode = _deferred_mode(meta, lib.p4c_mlp_wgrad_blocks(C.byref(desc)), nfl, device, branch)
    g.partials_mode = mode
    if mode != 0:
        g.wgrad_partials, g.wgrad_partials_floats = ptr(meta.partials[branch]), meta.partials[branch].numel()
        return rets
    for j, slot in enumerate(("g_w1", "g_b1", "g_w2", "g_b2", "g_gamma", "g_beta")):
        buf, ret = _grad_target(ws[j], needed[j])
        setattr(g, slot, ptr(buf))
        rets[j] = ret
    wsp = _wgrad_workspace((device, branch), nfl)
    g.wgrad_partials, g.wgrad_partials_floats = ptr(wsp), wsp.numel()
    return rets


def _check_hidden(t: torch.Tensor, what: str) -> None:
    if t.shape[-1] != HIDDEN:
        raise NotImplementedError(
            f"{what}: the fused kernels are built for hidden_dims={HIDDEN}, got {t.shape[-1]}"
        )


def padded_width(width: int, dtype: torch.dtype) -> int:
    """Row stride that lets the tensor-core engine fetch a `width`-wide row in 16-byte chunks."""
    return (width + 7) // 8 * 8 if dtype == torch.bfloat16 else width


def empty_rows(shape: Sequence[int], dtype: torch.dtype, device, zero: bool = False) -> torch.Tensor:
    """(…, R, W) tensor whose rows are padded to `padded_width`, returned as the (…, R, W) view:
    what the tensor-core MLP reads / writes without a copy.  The padding must hold finite values
    before the tensor is READ by a kernel (`zero=True`, or a producer kernel that writes it)."""
    *lead, w = shape
    ld = padded_width(w, dtype)
    if ld == w:
        return torch.empty(*lead, w, dtype=dtype, device=device)
    buf = (torch.zeros if zero else torch.empty)(*lead, ld, dtype=dtype, device=device)
    return buf[..., :w]


def _bwd_recomputes(lib, desc: Mlp) -> bool:
    """True when the backward of the problem `desc` (described as the BACKWARD sees it: plain sources)
    will be served by the pipelined engine that recomputes h1 / LayerNorm statistics
    (p4c_mlp_bwd_caps & 1): the forward then saves neither."""
    return bool(lib.p4c_mlp_bwd_caps(C.byref(desc)) & 1)


# -------------------------------------------------------------------- fused MLP
# ---- gradient hand-over between the consumers of one tensor ------------------------------------------------
# A representation that feeds two nodes of ours (grid embedding -> g2m senders + grid encoder; a mesh level ->
# senders of one InteractionNet + receivers of another) receives two input gradients that torch autograd sums
# with an element-wise pass.  The first backward to run registers the dense gradient tensor it returns, keyed by
# the forward tensor's storage; a later backward of the SAME autograd pass whose share comes out of an
# accumulating kernel (the by-sender segment-sum + GEMM) adds into that tensor and returns None instead.  Legal
# because the producer's backward cannot start before every consumer's has finished; entries die with the pass
# (torch's graph-task id), so repeated backward passes over recycled addresses never meet a stale tensor.
# OPT-IN per call site (`share_grads`): it is only correct when the tensor has exactly these two consumers -- with
# a third one autograd may already have summed the registered tensor into a fresh buffer before the hand-over.
_GRAD_REG: dict = {}
_GRAD_REG_TASK = [-1]
GRAD_HANDOVERS = [0]  # hand-overs performed (tests / diagnostics)


def _grad_key(t: torch.Tensor):
    return (t.data_ptr(), tuple(t.shape), tuple(t.stride()), t.dtype)


def _grad_task() -> int:
    task = torch._C._current_graph_task_id() if hasattr(torch._C, "_current_graph_task_id") else -1
    if task != _GRAD_REG_TASK[0]:
        _GRAD_REG.clear()
        _GRAD_REG_TASK[0] = task
    return task


def _grad_register(fwd_tensor: torch.Tensor, grad: Optional[torch.Tensor]) -> None:
    if grad is None or _grad_task() < 0:
        return
    if fwd_tensor.dim() == 3 and grad.is_contiguous() and tuple(grad.shape) == tuple(fwd_tensor.shape):
        _GRAD_REG[_grad_key(fwd_tensor)] = grad


def _grad_lookup(fwd_tensor: torch.Tensor) -> Optional[torch.Tensor]:
    if _grad_task() < 0:
        return None
    return _GRAD_REG.get(_grad_key(fwd_tensor))


class _MLPFn(torch.autograd.Function):
    """y = [LN](W2 silu(W1 x + b1) + b2) [+ x]   for one dense, identity-indexed source."""

    @staticmethod
    def forward(ctx, meta: Optional[MlpMeta], x, w1, b1, w2, b2, gamma, beta, residual: bool, share_grads: bool = False):
        lib = _lib.load()
        ctx.share_grads = share_grads
        shared = x.dim() == 2
        x = _as_rows(x)
        B = 1 if shared else x.shape[0]
        R, K = x.shape[-2], x.shape[-1]
        out_features = w2.shape[0]
        if w1.shape[0] != HIDDEN:
            raise NotImplementedError(f"fused MLP needs hidden width {HIDDEN}, got {w1.shape[0]}")
        if residual and (K != HIDDEN or out_features != HIDDEN):
            raise ValueError("residual needs 64-wide input and output")
        if residual and not x.is_contiguous():
            x = x.contiguous()
        dt = x.dtype
        need_grad = any(ctx.needs_input_grad)
        out = torch.empty((R, out_features) if shared else (B, R, out_features), dtype=dt, device=x.device)
        desc = Mlp()
        _fill_mlp(desc, R, B, 0, [_SrcSpec(x)], (w1, b1, w2, b2, gamma, beta), out_features, dt, meta)
        # the pipelined backward engine recomputes h1 and the LayerNorm statistics: nothing to save
        save = need_grad and not _bwd_recomputes(lib, desc)
        h1 = torch.empty(B * R, HIDDEN, dtype=dt, device=x.device) if save else None
        stats = torch.empty(B * R, 2, dtype=torch.float32, device=x.device) if (save and gamma is not None) else None
        check(
            lib.p4c_mlp_forward(C.byref(desc), ptr(x) if residual else None, ptr(out), None,
                                ptr(h1), ptr(stats), stream()),
            "p4c_mlp_forward",
        )
        ctx.save_for_backward(x, w1, b1, w2, b2, gamma, beta, h1, stats)
        ctx.residual = residual
        ctx.meta = meta
        ctx.branch = _BRANCH
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
        ws = (w1, b1, w2, b2, gamma, beta)
        desc = Mlp()
        _fill_mlp(desc, R, B, 0, [_SrcSpec(x)], ws, out_features, dt, ctx.meta)
        g = MlpBwd()
        g.dy_a, g.dy_a_batch_stride = ptr(dy), R * out_features
        g.h1, g.stats = ptr(h1), ptr(stats)
        dx = None
        if ctx.needs_input_grad[1]:
            dx = empty_rows(x.shape, dt, x.device)
            g.dx[0], g.dx_ld[0], g.dx_accumulate[0] = ptr(dx), _row_layout(dx)[0], 0
            g.dx0_add_dy = 2 if ctx.residual else 0
        elif ctx.residual:
            raise RuntimeError("residual MLP with a non-differentiable input is not supported")
        rets = _bind_wgrads(lib, desc, g, ws, ctx.needs_input_grad[2:8], ctx.meta, x.device, ctx.branch)
        check(lib.p4c_mlp_backward(C.byref(desc), C.byref(g), stream()), "p4c_mlp_backward")
        if ctx.share_grads and dx is not None and not shared and _row_layout(dx)[0] == K:
            _grad_register(x, dx)
        return (None, dx, *rets, None, None)


def fused_mlp(x: torch.Tensor, weights: Sequence[Optional[torch.Tensor]], residual: bool = False,
              meta: Optional[MlpMeta] = None, share_grads: bool = False) -> torch.Tensor:
    """share_grads: x has exactly one other consumer, an interaction_net(..., share_grads=True) that gathers it as
    its senders -- the two input gradients are summed by that call's by-sender kernel, not by autograd."""
    w1, b1, w2, b2, gamma, beta = weights
    return _MLPFn.apply(meta, x, w1, b1, w2, b2, gamma, beta, residual, share_grads)


# --------------------------------------------------------------- InteractionNet
def _rows(t: torch.Tensor) -> int:
    return t.shape[-2]


def seg_linear(src: torch.Tensor, csr: Csr, batch: int, w1: torch.Tensor, col0: int, out: torch.Tensor,
               accumulate: bool, mean: bool = False, meta: Optional[MlpMeta] = None) -> None:
    """out[b,r,:] (+)= (segment sum of src rows of segment r) @ w1[:, col0:col0+64]  (bf16, one kernel)."""
    packed = meta.image_ptr(src.dtype) if meta is not None else None
    check(
        _lib.load().p4c_seg_linear(
            ptr(src), src.shape[1] * HIDDEN, ptr(csr.rowptr), ptr(csr.perm), ptr(csr.inv_deg) if mean else None,
            csr.num_seg, batch, ptr(w1), w1.shape[1], col0, packed, col0 // HIDDEN, ptr(out),
            out.shape[1] * HIDDEN, int(accumulate), _seg_tile_rows(src.shape[1], csr.num_seg),
            dtype_code(src.dtype), stream(),
        ),
        "p4c_seg_linear",
    )


def _seg_part(src: torch.Tensor, csr: Csr, batch: int, col0: int, out: torch.Tensor, accumulate: bool) -> SegPart:
    q = SegPart()
    q.rowptr, q.perm, q.scale = ptr(csr.rowptr), ptr(csr.perm), None
    q.num_seg, q.batch, q.src_batch_stride = csr.num_seg, batch, src.shape[1] * HIDDEN
    q.col0, q.kblock = col0, col0 // HIDDEN
    q.out, q.out_batch_stride = ptr(out), out.shape[1] * HIDDEN
    q.accumulate, q.tile_rows = int(accumulate), _seg_tile_rows(src.shape[1], csr.num_seg)
    return q


def seg_linear_pair(src: torch.Tensor, w1: torch.Tensor, meta: Optional[MlpMeta], batch: int,
                    a: Tuple[Csr, int, torch.Tensor, bool], b: Tuple[Csr, int, torch.Tensor, bool]) -> None:
    """Two seg_linear problems (csr, col0, out, accumulate) over the same rows `src` in ONE launch.  When both
    write the same `out` (sender == receiver nodes) they become the two K-blocks of one product and `out` is
    updated once; otherwise the CTAs of the launch are split between them."""
    packed = meta.image_ptr(src.dtype) if meta is not None else None
    qa = _seg_part(src, a[0], batch, a[1], a[2], a[3])
    qb = _seg_part(src, b[0], batch, b[1], b[2], b[3])
    if a[2].data_ptr() == b[2].data_ptr():
        qb.tile_rows = qa.tile_rows = min(qa.tile_rows, qb.tile_rows)
    check(
        _lib.load().p4c_seg_linear_pair(ptr(src), ptr(w1), w1.shape[1], packed, C.byref(qa), C.byref(qb),
                                        dtype_code(src.dtype), stream()),
        "p4c_seg_linear_pair",
    )


def _batch_rows_csr(gi: GraphIndex, batch: int, device) -> Csr:
    """Segments {b * E + e : b < batch} over the rows of a (B, E, 64) tensor viewed as (B * E, 64):
    the batch sum of per-edge rows expressed as a segment sum (cached per edge set and batch)."""
    cache = gi.__dict__.setdefault("_batch_csr", {})
    key = (batch, str(device))
    if key not in cache:
        E = gi.num_edges
        rowptr = torch.arange(0, (E + 1) * batch, batch, dtype=torch.int32, device=device)
        perm = None
        if batch > 1:
            e = torch.arange(E, dtype=torch.int32, device=device)
            perm = (e.unsqueeze(1) + E * torch.arange(batch, dtype=torch.int32, device=device).unsqueeze(0))
            perm = perm.reshape(-1).contiguous()
        cache[key] = Csr(rowptr, perm, E, None)
    return cache[key]


class _InteractionNetFn(torch.autograd.Function):
    """
    (send, rec, edge) -> (rec_out, edge_out):
        msg      = edge_mlp(cat(edge, send[sender], rec[receiver]))
        aggr     = segment_{sum|mean}(msg by receiver)
        rec_out  = rec + aggr_mlp(cat(rec, aggr))
        edge_out = edge + msg                       (update_edges)
    Weights: 6 tensors per edge group followed by 6 per node chunk; `metas` (optional) holds the
    MlpMeta of each, in the same order.

    bf16 (tensor-core engine): 2 launches forward — the aggregation runs in the epilogue of the
    edge kernel when the edges are receiver-sorted with a uniform in-degree (m2g, up / down
    edges), otherwise in the prologue of the node kernel — and 4 backward (node MLP, edge MLP,
    and one fused segment-sum + GEMM per gathered source).  fp32 keeps separate segment sums.
    """

    @staticmethod
    def forward(ctx, gi: GraphIndex, update_edges: bool, mean: bool, metas, share_grads, send, rec, edge, *weights):
        lib = _lib.load()
        ctx.share_grads = bool(share_grads)
        for t, n in ((send, "send"), (rec, "rec"), (edge, "edge")):
            _check_hidden(t, f"InteractionNet {n} representation")
        same_nodes = (send.data_ptr() == rec.data_ptr() and send.shape == rec.shape
                      and send.stride() == rec.stride())
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
        metas = list(metas) if metas is not None else [None] * (n_eg + len(node_chunks))
        dev = rec.device
        need_grad = any(ctx.needs_input_grad)
        ctx.set_materialize_grads(False)
        tc = dt == torch.bfloat16 and bool(lib.p4c_get_engine())
        fuse_epilogue = tc and gi.uniform_deg > 0 and n_eg == 1

        msg = None if fuse_epilogue else torch.empty(B, E, HIDDEN, dtype=dt, device=dev)
        aggr = torch.empty(B, Nr, HIDDEN, dtype=dt, device=dev)
        edge_out = torch.empty(B, E, HIDDEN, dtype=dt, device=dev) if update_edges else None
        edge_descs = []
        for gidx, (grp, w) in enumerate(zip(gi.groups, ew)):
            n = grp.stop - grp.start
            if n == 0:
                edge_descs.append(None)
                continue
            desc = Mlp()
            _fill_mlp(
                desc, n, B, E,
                [
                    _SrcSpec(edge, row_offset=grp.start),
                    _SrcSpec(send, index=gi.send_idx, index_offset=grp.start),
                    _SrcSpec(rec, index=gi.rec_idx, index_offset=grp.start),
                ],
                w, HIDDEN, dt, metas[gidx],
            )
            edge_descs.append(desc)
        # h1 / LayerNorm statistics are saved only for the engines that read them back: the pipelined
        # backward recomputes both (mean aggregation scales dy per row, which that engine does not take)
        save_e = need_grad and not (
            tc and not mean and all(d is None or _bwd_recomputes(lib, d) for d in edge_descs))
        h1_e = torch.empty(B, E, HIDDEN, dtype=dt, device=dev) if save_e else None
        st_e = torch.empty(B, E, 2, dtype=torch.float32, device=dev) if save_e else None
        for gidx, (grp, w) in enumerate(zip(gi.groups, ew)):
            n = grp.stop - grp.start
            if n == 0:
                continue
            desc = edge_descs[gidx]
            o = grp.start * HIDDEN
            args = (
                _off(edge_out, o) if update_edges else None,
                _off(h1_e, o) if save_e else None,
                _off(st_e, grp.start * 2) if save_e else None,
            )
            if fuse_epilogue:
                check(
                    lib.p4c_mlp_forward_aggregate(
                        C.byref(desc), None, *args, ptr(aggr), gi.uniform_deg,
                        1.0 / gi.uniform_deg if mean else 1.0, stream(),
                    ),
                    "p4c_mlp_forward_aggregate(edge)",
                )
            else:
                check(lib.p4c_mlp_forward(C.byref(desc), None, _off(msg, o), *args, stream()),
                      "p4c_mlp_forward(edge)")
        seg_in_prologue = tc and not fuse_epilogue
        if not tc:
            segment_sum(msg, gi.by_rec, B, mean=mean, out=aggr)

        rec_b = rec if rec.dim() == 3 else rec.unsqueeze(0).expand(B, Nr, HIDDEN).contiguous()
        rec_out = torch.empty(B, Nr, HIDDEN, dtype=dt, device=dev)
        save_n = need_grad
        if tc and need_grad:
            save_n = False
            for ci, ((n0, n1), w) in enumerate(zip(node_chunks, nw)):
                bdesc = Mlp()  # the problem as the backward describes it (saved sums as plain rows)
                _fill_mlp(bdesc, n1 - n0, B, Nr, [_SrcSpec(rec_b, row_offset=n0), _SrcSpec(aggr, row_offset=n0)], w,
                          HIDDEN, dt, metas[n_eg + ci])
                save_n = save_n or not _bwd_recomputes(lib, bdesc)
        h1_n = torch.empty(B, Nr, HIDDEN, dtype=dt, device=dev) if save_n else None
        st_n = torch.empty(B, Nr, 2, dtype=torch.float32, device=dev) if save_n else None
        for ci, ((n0, n1), w) in enumerate(zip(node_chunks, nw)):
            desc = Mlp()
            if seg_in_prologue:
                aggr_src = _SrcSpec(msg, seg=gi.by_rec, seg_row_offset=n0, seg_mean=mean, seg_out=aggr)
            else:
                aggr_src = _SrcSpec(aggr, row_offset=n0)
            _fill_mlp(desc, n1 - n0, B, Nr, [_SrcSpec(rec_b, row_offset=n0), aggr_src], w, HIDDEN, dt,
                      metas[n_eg + ci])
            o = n0 * HIDDEN
            check(
                lib.p4c_mlp_forward(
                    C.byref(desc), _off(rec_b, o), _off(rec_out, o), None,
                    _off(h1_n, o) if save_n else None,
                    _off(st_n, n0 * 2) if save_n else None, stream(),
                ),
                "p4c_mlp_forward(node)",
            )
        del msg
        ctx.gi, ctx.update_edges, ctx.mean, ctx.B = gi, update_edges, mean, B
        ctx.rec_shared = rec.dim() == 2
        ctx.same_nodes = same_nodes and rec.dim() == 3
        ctx.metas, ctx.tc = metas, tc
        ctx.branch = _BRANCH
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
        metas, tc = ctx.metas, ctx.tc
        need_send, need_rec, need_edge = ctx.needs_input_grad[5:8]
        wneed = ctx.needs_input_grad[8:]
        same_nodes = ctx.same_nodes and need_send and need_rec

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

        # ---- node MLP backward: d_rec (incl. residual pass-through) and d_aggr
        d_rec = torch.empty(B, Nr, HIDDEN, dtype=dt, device=dev)
        d_aggr = torch.empty(B, Nr, HIDDEN, dtype=dt, device=dev)
        for ci, ((n0, n1), w) in enumerate(zip(node_chunks, nw)):
            desc = Mlp()
            _fill_mlp(
                desc, n1 - n0, B, Nr,
                [_SrcSpec(rec_b, row_offset=n0), _SrcSpec(aggr, row_offset=n0)], w, HIDDEN, dt, metas[n_eg + ci],
            )
            o = n0 * HIDDEN
            g = MlpBwd()
            g.dy_a, g.dy_a_batch_stride = _off(d_rec_out, o), Nr * HIDDEN
            if h1_n is not None:
                g.h1, g.stats = _off(h1_n, o), _off(st_n, n0 * 2)
            g.dx[0], g.dx_ld[0], g.dx_accumulate[0] = _off(d_rec, o), HIDDEN, 0
            g.dx[1], g.dx_ld[1], g.dx_accumulate[1] = _off(d_aggr, o), HIDDEN, 0
            g.dx0_add_dy = 2
            base = 6 * (n_eg + ci)
            wgrads[base: base + 6] = _bind_wgrads(lib, desc, g, w, wneed[base: base + 6], metas[n_eg + ci], dev,
                                                  ctx.branch)
            check(lib.p4c_mlp_backward(C.byref(desc), C.byref(g), stream()), "p4c_mlp_backward(node)")

        # ---- edge MLP backward
        dh1 = torch.empty(B, E, HIDDEN, dtype=dt, device=dev)
        # Edge representation shared by the batch (static embeddings): its gradient is
        # (sum_b dh1[b]) x W1[:, edge block] by linearity -> one fused segment-sum + GEMM over the
        # batch instead of a (B, E, 64) input gradient from the edge kernel and a batch sum.
        edge_by_linearity = need_edge and tc and edge.dim() == 2 and d_edge_out is None
        d_edge = None
        if need_edge:
            d_edge = (torch.empty(E, HIDDEN, dtype=dt, device=dev) if edge_by_linearity
                      else torch.empty(B, E, HIDDEN, dtype=dt, device=dev))
        d_send = None
        send_handed_over = False
        if need_send and not same_nodes:
            # another consumer of `send` already returned its share of the gradient in this pass: the by-sender
            # pass adds into it (see the hand-over notes above _MLPFn)
            held = _grad_lookup(send) if (ctx.share_grads and tc and send.dim() == 3) else None
            if held is not None and held.dtype == dt and tuple(held.shape) == (B, Ns, HIDDEN):
                d_send, send_handed_over = held, True
                GRAD_HANDOVERS[0] += 1
            else:
                d_send = torch.empty(B, Ns, HIDDEN, dtype=dt, device=dev)
        first_send = not send_handed_over
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
                w, HIDDEN, dt, metas[gidx],
            )
            o = grp.start * HIDDEN
            g = MlpBwd()
            g.dy_a = ptr(d_aggr)
            g.dy_a_index = _off(gi.rec_idx, grp.start)
            g.dy_a_scale = ptr(gi.by_rec.inv_deg) if ctx.mean else None
            g.dy_a_batch_stride = Nr * HIDDEN
            g.dy_b = _off(d_edge_out, o) if d_edge_out is not None else None
            if h1_e is not None:
                g.h1, g.stats = _off(h1_e, o), _off(st_e, grp.start * 2)
            g.dh1_out = _off(dh1, o)
            # receiver-sorted edges of uniform in-degree: the pipelined engine sums the receiver gradient
            # over each group of rows in its epilogue and accumulates it into d_rec (no pass over dh1)
            caps = lib.p4c_mlp_bwd_caps(C.byref(desc)) if (tc and not ctx.mean) else 0
            rec_in_epilogue = tc and need_rec and gi.uniform_deg > 0 and n_eg == 1 and bool(caps & 4)
            # static edge embedding shared by the batch: the pipelined engine accumulates dh1 x W1[:, edge block] over
            # the samples of a row block in TMEM and stores the (E, 64) gradient itself (no pass over dh1)
            edge_in_kernel = edge_by_linearity and bool(caps & 2)
            if edge_in_kernel:
                g.dx0_batch_sum = _off(d_edge, o)
            if rec_in_epilogue:
                g.dx_group_out, g.dx_group, g.dx_group_src = ptr(d_rec), gi.uniform_deg, 2
            if need_edge and not edge_by_linearity:
                g.dx[0], g.dx_ld[0], g.dx_accumulate[0] = _off(d_edge, o), HIDDEN, 0
                g.dx0_add_dy = 1 if d_edge_out is not None else 0
            wgrads[6 * gidx: 6 * gidx + 6] = _bind_wgrads(lib, desc, g, w, wneed[6 * gidx: 6 * gidx + 6],
                                                          metas[gidx], dev, ctx.branch)
            check(lib.p4c_mlp_backward(C.byref(desc), C.byref(g), stream()), "p4c_mlp_backward(edge)")
            # sender / receiver gradients by linearity: (segment-sum of dh1) x W1 slice
            if tc:
                if edge_by_linearity and not edge_in_kernel:
                    bc = _batch_rows_csr(gi, B, dev)
                    seg_linear(dh1.view(1, B * E, HIDDEN), Csr(bc.rowptr[grp.start: grp.stop + 1], bc.perm, n),
                               1, w[0], 0, d_edge[grp.start: grp.stop].unsqueeze(0), False, meta=metas[gidx])
                rec_pass = need_rec and not rec_in_epilogue
                if rec_pass and need_send:
                    # receiver and sender gradient of the same dh1 in one launch (one product over two
                    # K-blocks and one update of d_rec when the senders are the receivers)
                    send_part = ((grp.by_send, HIDDEN, d_rec, True) if same_nodes
                                 else (grp.by_send, HIDDEN, d_send, not first_send))
                    seg_linear_pair(dh1, w[0], metas[gidx], B, (grp.by_rec, 2 * HIDDEN, d_rec, True), send_part)
                    first_send = False
                else:
                    if rec_pass:
                        seg_linear(dh1, grp.by_rec, B, w[0], 2 * HIDDEN, d_rec, True, meta=metas[gidx])
                    if need_send:
                        if same_nodes:
                            seg_linear(dh1, grp.by_send, B, w[0], HIDDEN, d_rec, True, meta=metas[gidx])
                        else:
                            seg_linear(dh1, grp.by_send, B, w[0], HIDDEN, d_send, not first_send, meta=metas[gidx])
                        first_send = False
            else:
                if need_rec:
                    s_r = segment_sum(dh1, grp.by_rec, B)
                    rows_matmul(s_r, w[0], 2 * HIDDEN, d_rec, accumulate=True)
                if need_send:
                    s_s = segment_sum(dh1, grp.by_send, B)
                    if same_nodes:
                        rows_matmul(s_s, w[0], HIDDEN, d_rec, accumulate=True)
                    else:
                        rows_matmul(s_s, w[0], HIDDEN, d_send, accumulate=not first_send)
                    first_send = False
        if d_send is not None and first_send:
            d_send.zero_()

        g_send = None
        if need_send and not same_nodes and not send_handed_over:
            g_send = batch_sum(d_send) if send.dim() == 2 else d_send
            if ctx.share_grads and tc and send.dim() == 3:
                _grad_register(send, g_send)
        g_rec = (batch_sum(d_rec) if ctx.rec_shared else d_rec) if need_rec else None
        if ctx.share_grads and tc and g_rec is not None and not ctx.rec_shared:
            _grad_register(rec_b, g_rec)
        if edge_by_linearity:
            g_edge = d_edge
        else:
            g_edge = (batch_sum(d_edge) if edge.dim() == 2 else d_edge) if need_edge else None
        return (None, None, None, None, None, g_send, g_rec, g_edge, *wgrads)


def interaction_net(gi: GraphIndex, send, rec, edge, weights: Sequence[torch.Tensor],
                    update_edges: bool, mean: bool, metas: Optional[Sequence[Optional[MlpMeta]]] = None,
                    share_grads: bool = False):
    """share_grads: the caller guarantees that `send` and `rec` (when they need gradients) have at most ONE other
    consumer each, which is also called with share_grads: the second of the two backward passes adds its share
    into the first one's tensor (by-sender kernel) instead of leaving the sum to autograd."""
    return _InteractionNetFn.apply(gi, update_edges, mean, metas, share_grads, send, rec, edge, *weights)
