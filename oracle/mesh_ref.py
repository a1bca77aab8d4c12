"""
ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see oracle/graph_ref.py header).

Independent restatement of the neural-lam `create_mesh.py` recipe (the algorithm mfai v5.0.1
ports) with scipy's cKDTree for the neighbour queries, exactly as the
upstream does (kd-tree ball query of radius 0.67*dm for g2m, 4-NN for m2g, 1-NN between levels).
Used to cross-check py4cast_b200/graph_build.py, which replaces the kd-tree by exact windowed
searches with explicit tie-breaking.  Returns plain python sets of (sender, receiver) pairs and
per-receiver sorted distance lists, so that ties (equidistant mesh nodes) can be compared by
distance rather than by index.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree


def level_nodes(xy: np.ndarray, n: int) -> np.ndarray:
    xm, xM = np.amin(xy[0][0, :]), np.amax(xy[0][0, :])
    ym, yM = np.amin(xy[1][:, 0]), np.amax(xy[1][:, 0])
    dx, dy = (xM - xm) / n, (yM - ym) / n
    lx = np.linspace(xm + dx / 2, xM - dx / 2, n)
    ly = np.linspace(ym + dy / 2, yM - dy / 2, n)
    mg = np.meshgrid(lx, ly)
    return np.stack([mg[0].reshape(-1), mg[1].reshape(-1)], axis=1)


def mesh_levels(xy: np.ndarray, n_max_levels=None):
    nx = 3
    nlev = int(np.log(max(xy.shape)) / np.log(nx))
    nleaf = nx**nlev
    levels = nlev - 1
    if n_max_levels:
        levels = min(levels, n_max_levels)
    return [int(nleaf / (nx**lev)) for lev in range(1, levels + 1)]


def lattice_edge_set(n: int):
    """undirected 4-neighbour grid + both diagonals, made directed both ways"""
    edges = set()
    for y in range(n):
        for x in range(n):
            u = y * n + x
            for dy_, dx_ in ((0, 1), (1, 0), (1, 1), (1, -1)):
                yy, xx = y + dy_, x + dx_
                if 0 <= yy < n and 0 <= xx < n:
                    v = yy * n + xx
                    edges.add((u, v))
                    edges.add((v, u))
    return edges


def g2m_edge_set(xy: np.ndarray, n: int):
    grid = np.stack([xy[0].reshape(-1), xy[1].reshape(-1)], axis=1)
    mesh = level_nodes(xy, n)
    dm = np.sqrt(np.sum((mesh[1] - mesh[0]) ** 2))
    tree = cKDTree(grid)
    out = set()
    for m, pos in enumerate(mesh):
        for g in tree.query_ball_point(pos, dm * 0.67):
            out.add((int(g), m))
    return out


def m2g_sorted_distances(xy: np.ndarray, n: int):
    """per grid node: sorted distances of its 4 nearest mesh nodes, and their indices"""
    grid = np.stack([xy[0].reshape(-1), xy[1].reshape(-1)], axis=1)
    mesh = level_nodes(xy, n)
    d, idx = cKDTree(mesh).query(grid, 4)
    return d, idx


def down_edge_set(xy: np.ndarray, n_fine: int, n_coarse: int):
    fine, coarse = level_nodes(xy, n_fine), level_nodes(xy, n_coarse)
    _, idx = cKDTree(coarse).query(fine, 1)
    return {(int(u), v) for v, u in enumerate(idx)}
