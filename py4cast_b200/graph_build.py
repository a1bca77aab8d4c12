"""
Deterministic mesh-graph construction for GraphLAM / HiLAM / HiLAMParallel.

Host-side, set-up-time code (numpy): runs once per model, never on the AR
training path.  It is the *definition* of edge indexing for this package —
the recipe follows the neural-lam `create_mesh.py` algorithm that
`mfai.pytorch.models` (v5.0.1, requirements.txt:26, source absent from
/root/reference) ports; SURVEY.md Appendix A.8 restates it.  Reference call
sites: py4cast/lightning.py:141-144,271 (`rank_zero_setup(settings, meshgrid)`),
tests/test_models.py:44-60,76-79 (meshgrid format `(2, H, W)`).

Choices fixed here because the upstream ordering is an artefact of networkx
iteration order (unpinned):
  * node order: row-major `(iy, ix)` per level; grid nodes row-major `(i, j)`;
  * edge order: sorted by (receiver, sender) — i.e. edges are stored
    receiver-sorted, so the CSR row pointer is a plain histogram prefix sum;
  * k-NN / ball-query ties are broken towards the lowest node index
    (distances in float64), not by a KD-tree's traversal order;
  * `up` edges are the `down` edges with sender/receiver swapped and the SAME
    features (neural-lam clones the down graph and swaps the index rows).
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import torch

MESH_REFINEMENT = 3  # neural-lam `nx`
G2M_RADIUS_SCALE = 0.67  # neural-lam DM_SCALE


@dataclass
class EdgeSet:
    """Directed edges sender -> receiver with 0-based LOCAL node ids."""

    edge_index: torch.Tensor  # (2, E) int64; row 0 = sender, row 1 = receiver
    features: torch.Tensor  # (E, 3) float32: (len, vdiff_x, vdiff_y), normalised
    num_send: int
    num_rec: int

    @property
    def num_edges(self) -> int:
        return int(self.edge_index.shape[1])


@dataclass
class MeshGraph:
    hierarchical: bool
    grid_shape: tuple
    level_sizes: List[int]  # nodes per side, finest first
    mesh_static_features: List[torch.Tensor]  # per level (N_l, 2); 1 entry if flat
    m2m: List[EdgeSet]  # per level (same-level edges); 1 entry if flat (merged)
    g2m: EdgeSet = None
    m2g: EdgeSet = None
    mesh_up: List[EdgeSet] = field(default_factory=list)  # level l -> l+1
    mesh_down: List[EdgeSet] = field(default_factory=list)  # level l+1 -> l

    @property
    def num_levels(self) -> int:
        return len(self.m2m)

    @property
    def num_grid_nodes(self) -> int:
        return self.grid_shape[0] * self.grid_shape[1]

    def num_mesh_nodes(self, level: int = 0) -> int:
        return int(self.mesh_static_features[level].shape[0])


# ----------------------------------------------------------------- helpers
def _canonical(send, rec, feats):
    """Sort edges by (receiver, sender)."""
    order = np.lexsort((send, rec))
    return send[order], rec[order], feats[order]


def _edge_features(pos_send, pos_rec):
    vdiff = pos_send - pos_rec
    length = np.sqrt(np.sum(vdiff * vdiff, axis=1))
    return np.concatenate([length[:, None], vdiff], axis=1)


def _level_positions(xy: np.ndarray, n: int):
    """neural-lam mk_2d_graph: n x n nodes at cell centres of the bounding box."""
    xm, xM = np.amin(xy[0][0, :]), np.amax(xy[0][0, :])
    ym, yM = np.amin(xy[1][:, 0]), np.amax(xy[1][:, 0])
    dx = (xM - xm) / n
    dy = (yM - ym) / n
    lx = np.linspace(xm + dx / 2, xM - dx / 2, n)
    ly = np.linspace(ym + dy / 2, yM - dy / 2, n)
    gx, gy = np.meshgrid(lx, ly)  # [iy, ix]
    pos = np.stack([gx.reshape(-1), gy.reshape(-1)], axis=1)  # node id = iy*n + ix
    return pos, lx, ly


def _lattice_edges(n: int):
    """8-neighbour directed lattice on n x n nodes (4-neighbour + both diagonals)."""
    iy, ix = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    iy = iy.reshape(-1)
    ix = ix.reshape(-1)
    send, rec = [], []
    for oy in (-1, 0, 1):
        for ox in (-1, 0, 1):
            if oy == 0 and ox == 0:
                continue
            jy, jx = iy + oy, ix + ox
            ok = (jy >= 0) & (jy < n) & (jx >= 0) & (jx < n)
            send.append((jy * n + jx)[ok])
            rec.append((iy * n + ix)[ok])
    return np.concatenate(send), np.concatenate(rec)


def _nearest_window(pos_q, lx, ly, k, radius=None):
    """
    For every query point, candidate lattice nodes in a window around the cell
    containing it; returns either the k nearest (ties -> lowest node id) or all
    nodes within `radius` (d^2 <= r^2).  Exact float64, KD-tree free.
    Lattice node id = iy*len(lx) + ix at position (lx[ix], ly[iy]).
    """
    nx_, ny_ = len(lx), len(ly)
    dx = (lx[1] - lx[0]) if nx_ > 1 else 1.0
    dy = (ly[1] - ly[0]) if ny_ > 1 else 1.0
    reach = radius if radius is not None else np.sqrt(dx * dx + dy * dy)
    hx = int(np.ceil(reach / dx)) + 1
    hy = int(np.ceil(reach / dy)) + 1
    cx = np.clip(np.rint((pos_q[:, 0] - lx[0]) / dx).astype(np.int64), 0, nx_ - 1)
    cy = np.clip(np.rint((pos_q[:, 1] - ly[0]) / dy).astype(np.int64), 0, ny_ - 1)
    offs = [(oy, ox) for oy in range(-hy, hy + 1) for ox in range(-hx, hx + 1)]
    nq = pos_q.shape[0]
    cand = np.empty((nq, len(offs)), dtype=np.int64)
    d2 = np.empty((nq, len(offs)), dtype=np.float64)
    for c, (oy, ox) in enumerate(offs):
        jy, jx = cy + oy, cx + ox
        ok = (jy >= 0) & (jy < ny_) & (jx >= 0) & (jx < nx_)
        jyc, jxc = np.clip(jy, 0, ny_ - 1), np.clip(jx, 0, nx_ - 1)
        ddx = pos_q[:, 0] - lx[jxc]
        ddy = pos_q[:, 1] - ly[jyc]
        dd = ddx * ddx + ddy * ddy
        d2[:, c] = np.where(ok, dd, np.inf)
        cand[:, c] = np.where(ok, jyc * nx_ + jxc, np.iinfo(np.int64).max)
    if radius is not None:
        inside = d2 <= radius * radius
        q_idx, c_idx = np.nonzero(inside)
        return q_idx, cand[q_idx, c_idx]
    # k nearest with (d2, id) lexicographic order
    order = np.lexsort((cand, d2), axis=1)[:, :k]
    rows = np.arange(nq)[:, None]
    return np.repeat(np.arange(nq), k), cand[rows, order].reshape(-1)


# ------------------------------------------------------------------ builder
def build_mesh_graph(
    meshgrid,
    hierarchical: bool,
    n_max_levels: Optional[int] = None,
) -> MeshGraph:
    """
    meshgrid: (2, H, W) array/tensor; [0] = x varying along axis 1, [1] = y along axis 0
    (reference Statics.meshgrid, base.py:216-230; tests/test_models.py:54-60).
    """
    xy = np.asarray(meshgrid.cpu() if isinstance(meshgrid, torch.Tensor) else meshgrid)
    xy = xy.astype(np.float64)
    if xy.ndim != 3 or xy.shape[0] != 2:
        raise ValueError(f"meshgrid must have shape (2, H, W), got {xy.shape}")
    H, W = xy.shape[1:]
    nx = MESH_REFINEMENT
    nlev = int(np.log(max(xy.shape)) / np.log(nx))
    nleaf = nx**nlev
    mesh_levels = nlev - 1
    if n_max_levels:
        mesh_levels = min(mesh_levels, int(n_max_levels))
    if mesh_levels < 1 or (hierarchical and mesh_levels < 2):
        raise ValueError(
            f"grid {H}x{W} is too small for a {'hierarchical ' if hierarchical else ''}"
            f"mesh (levels={mesh_levels})"
        )
    sizes = [nleaf // nx**lev for lev in range(1, mesh_levels + 1)]

    level_pos, level_axes, level_edges = [], [], []
    for n in sizes:
        pos, lx, ly = _level_positions(xy, n)
        s, r = _lattice_edges(n)
        level_pos.append(pos)
        level_axes.append((lx, ly))
        level_edges.append((s, r, _edge_features(pos[s], pos[r])))

    grid_pos = np.stack([xy[0].reshape(-1), xy[1].reshape(-1)], axis=1)  # id = i*W + j
    pos_max = np.max(np.abs(xy))

    raw = {}
    if hierarchical:
        raw["m2m"] = [
            _canonical(s, r, f) + (len(level_pos[l]), len(level_pos[l]))
            for l, (s, r, f) in enumerate(level_edges)
        ]
        raw["down"], raw["up"] = [], []
        for l in range(mesh_levels - 1):
            fine, coarse = level_pos[l], level_pos[l + 1]
            lx, ly = level_axes[l + 1]
            q, u = _nearest_window(fine, lx, ly, k=1)  # fine node q <- coarse node u
            f = _edge_features(coarse[u], fine[q])
            ds, dr, df = _canonical(u, q, f)
            raw["down"].append((ds, dr, df, len(coarse), len(fine)))
            us, ur, uf = _canonical(dr.copy(), ds.copy(), df.copy())
            raw["up"].append((us, ur, uf, len(fine), len(coarse)))
        mesh_feats = [p / pos_max for p in level_pos]
    else:
        # merged multiscale mesh: coarse node (iy,ix) of level l coincides with
        # node (3*iy+1, 3*ix+1) of level l-1
        send_all, rec_all, feat_all = [], [], []
        for l, (s, r, f) in enumerate(level_edges):
            n = sizes[l]
            iy, ix = np.divmod(np.arange(n * n), n)
            for _ in range(l):
                iy, ix = nx * iy + 1, nx * ix + 1
            to_fine = iy * sizes[0] + ix
            send_all.append(to_fine[s])
            rec_all.append(to_fine[r])
            feat_all.append(f)
        s, r, f = _canonical(
            np.concatenate(send_all), np.concatenate(rec_all), np.concatenate(feat_all)
        )
        raw["m2m"] = [(s, r, f, len(level_pos[0]), len(level_pos[0]))]
        mesh_feats = [level_pos[0] / pos_max]

    # grid <-> bottom mesh
    bottom = level_pos[0]
    lx, ly = level_axes[0]
    dm = np.sqrt(np.sum((bottom[1] - bottom[0]) ** 2))
    g, m = _nearest_window(grid_pos, lx, ly, k=0, radius=dm * G2M_RADIUS_SCALE)
    raw["g2m"] = _canonical(g, m, _edge_features(grid_pos[g], bottom[m])) + (
        len(grid_pos),
        len(bottom),
    )
    g, m = _nearest_window(grid_pos, lx, ly, k=4)
    raw["m2g"] = _canonical(m, g, _edge_features(bottom[m], grid_pos[g])) + (
        len(bottom),
        len(grid_pos),
    )

    longest_edge = max(np.max(e[2][:, 0]) for e in raw["m2m"])

    def mk(e) -> EdgeSet:
        s, r, f, ns, nr = e
        return EdgeSet(
            edge_index=torch.from_numpy(np.stack([s, r]).astype(np.int64)),
            features=torch.from_numpy((f / longest_edge).astype(np.float32)),
            num_send=int(ns),
            num_rec=int(nr),
        )

    return MeshGraph(
        hierarchical=hierarchical,
        grid_shape=(H, W),
        level_sizes=sizes if hierarchical else sizes[:1],
        mesh_static_features=[torch.from_numpy(p.astype(np.float32)) for p in mesh_feats],
        m2m=[mk(e) for e in raw["m2m"]],
        g2m=mk(raw["g2m"]),
        m2g=mk(raw["m2g"]),
        mesh_up=[mk(e) for e in raw.get("up", [])],
        mesh_down=[mk(e) for e in raw.get("down", [])],
    )


def unit_meshgrid(H: int, W: int) -> torch.Tensor:
    """(2,H,W) meshgrid with both coordinates normalised to [0,1]
    (reference access.py:272-279 normalisation of Statics.meshgrid)."""
    x = torch.linspace(0.0, 1.0, W, dtype=torch.float64)
    y = torch.linspace(0.0, 1.0, H, dtype=torch.float64)
    yy, xx = torch.meshgrid(y, x, indexing="ij")
    return torch.stack([xx, yy]).to(torch.float32)
