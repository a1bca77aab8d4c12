"""Mesh-graph builder: sizes pinned by SURVEY.md §8, canonical ordering, and cross-check of the
edge sets against the kd-tree restatement of the upstream recipe (oracle/mesh_ref.py)."""
import numpy as np
import pytest
import torch

from helpers import pixel_meshgrid
from oracle import mesh_ref
from py4cast_b200.graph_build import build_mesh_graph, unit_meshgrid


def _pairs(es):
    return set(zip(es.edge_index[0].tolist(), es.edge_index[1].tolist()))


def test_sizes_64x64_match_survey():
    g = build_mesh_graph(pixel_meshgrid(64, 64), hierarchical=True)
    assert g.level_sizes == [9, 3]
    assert [e.num_edges for e in g.m2m] == [544, 40]
    assert g.g2m.num_edges == 5436 and g.m2g.num_edges == 16384
    assert [e.num_edges for e in g.mesh_up] == [81] and [e.num_edges for e in g.mesh_down] == [81]
    deg = torch.bincount(g.g2m.edge_index[1])
    assert 64 <= int(deg.min()) and int(deg.max()) <= 68


def test_sizes_titan_match_survey():
    g = build_mesh_graph(unit_meshgrid(512, 640), hierarchical=True, n_max_levels=3)
    assert g.level_sizes == [81, 27, 9]
    assert [e.num_edges for e in g.m2m] == [51520, 5512, 544]
    assert g.g2m.num_edges == 459668 and g.m2g.num_edges == 1310720
    assert [e.num_edges for e in g.mesh_up] == [6561, 729]
    flat = build_mesh_graph(unit_meshgrid(512, 640), hierarchical=False)
    assert flat.m2m[0].num_edges == 57616 and flat.num_mesh_nodes() == 6561


@pytest.mark.parametrize("hier", [False, True])
def test_edges_are_receiver_sorted_and_local(hier):
    g = build_mesh_graph(pixel_meshgrid(40, 64), hierarchical=hier)
    sets = [g.g2m, g.m2g] + g.m2m + g.mesh_up + g.mesh_down
    for es in sets:
        s, r = es.edge_index
        assert int(s.min()) >= 0 and int(s.max()) < es.num_send
        assert int(r.min()) >= 0 and int(r.max()) == es.num_rec - 1
        key = r * es.num_send + s
        assert bool((key[1:] > key[:-1]).all()), "edges must be strictly sorted by (receiver, sender)"
        assert es.features.shape == (es.num_edges, 3) and es.features.dtype == torch.float32
    # every grid node receives exactly 4 mesh edges
    assert bool((torch.bincount(g.m2g.edge_index[1]) == 4).all())


@pytest.mark.parametrize("shape", [(64, 64), (40, 64)])
def test_edge_sets_match_kdtree_restatement(shape):
    H, W = shape
    mg = pixel_meshgrid(H, W)
    xy = mg.numpy().astype(np.float64)
    g = build_mesh_graph(mg, hierarchical=True)
    sizes = mesh_ref.mesh_levels(xy)
    assert sizes == g.level_sizes
    for l, n in enumerate(sizes):
        assert _pairs(g.m2m[l]) == mesh_ref.lattice_edge_set(n)
    # ball query: identical except for pairs exactly on the radius (kd-tree rounding)
    ref = mesh_ref.g2m_edge_set(xy, sizes[0])
    ours = _pairs(g.g2m)
    assert len(ref ^ ours) <= 0.001 * len(ref)
    # 4-NN: same multiset of distances per grid node (ties may pick different equidistant nodes)
    d_ref, idx_ref = mesh_ref.m2g_sorted_distances(xy, sizes[0])
    s, r = g.m2g.edge_index
    mesh = mesh_ref.level_nodes(xy, sizes[0])
    grid = np.stack([xy[0].reshape(-1), xy[1].reshape(-1)], axis=1)
    d_ours = np.sqrt(((mesh[s.numpy()] - grid[r.numpy()]) ** 2).sum(1)).reshape(-1, 4)
    np.testing.assert_allclose(np.sort(d_ours, axis=1), np.sort(d_ref, axis=1), rtol=0, atol=1e-9)
    for l in range(len(sizes) - 1):
        assert _pairs(g.mesh_down[l]) == mesh_ref.down_edge_set(xy, sizes[l], sizes[l + 1])
        up = g.mesh_up[l].edge_index
        assert set(zip(up[1].tolist(), up[0].tolist())) == _pairs(g.mesh_down[l])


def test_flat_mesh_is_union_of_levels():
    mg = pixel_meshgrid(64, 64)
    hi = build_mesh_graph(mg, hierarchical=True)
    flat = build_mesh_graph(mg, hierarchical=False)
    assert flat.m2m[0].num_edges == sum(e.num_edges for e in hi.m2m)
    assert flat.num_mesh_nodes() == hi.num_mesh_nodes(0)
    # features normalised by the longest m2m edge
    assert abs(float(flat.m2m[0].features[:, 0].max()) - 1.0) < 1e-6


def test_determinism_and_errors():
    a = build_mesh_graph(pixel_meshgrid(64, 64), True)
    b = build_mesh_graph(pixel_meshgrid(64, 64), True)
    assert torch.equal(a.g2m.edge_index, b.g2m.edge_index) and torch.equal(a.m2g.features, b.m2g.features)
    with pytest.raises(ValueError):
        build_mesh_graph(pixel_meshgrid(8, 8), True)
    with pytest.raises(ValueError):
        build_mesh_graph(torch.zeros(3, 4, 4), False)
