"""
ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under py4cast_b200/ may import this.

PARITY UNPINNED for this file: the graph-model family the reference trains
(GraphLAM / HiLAM / HiLAMParallel / InteractionNet / make_mlp) lives in the
third-party package `mfai` (pinned `mfai @ git+https://github.com/meteofrance/mfai@v5.0.1`,
reference requirements.txt:26; together with torch-geometric==2.6.1,
requirements.txt:23), which is NOT present under /root/reference nor
installable here.  This is a plain-PyTorch (index_select + index_add_)
restatement of the published neural-lam algorithm that mfai ports
(SURVEY.md §3.2 and Appendix A.1-A.7), anchored on the reference's own call
sites: lightning.py:265-279 (construction), :591-596 (forward (B,N,F_in) ->
(B,N,F_out)), tests/test_models.py:63-142 (kwargs, rank_zero_setup, flatten
convention) and config/CLI/model/hilam.yaml:18-25 (settings).
No golden vectors exist in the reference for it.

State-dict key names are shared with py4cast_b200.nlam so the same weights
can be loaded in both.
"""

from __future__ import annotations

from typing import List, Optional, Sequence

import torch
from torch import nn


def make_mlp(blueprint: Sequence[int], layer_norm: bool = True) -> nn.Sequential:
    """A.1: Linear, then (SiLU, Linear) per remaining pair, optional trailing LayerNorm."""
    layers: List[nn.Module] = []
    for i, (a, b) in enumerate(zip(blueprint[:-1], blueprint[1:])):
        if i > 0:
            layers.append(nn.SiLU())
        layers.append(nn.Linear(a, b))
    if layer_norm:
        layers.append(nn.LayerNorm(blueprint[-1]))
    return nn.Sequential(*layers)


class SplitMLPs(nn.Module):
    """A.2: one MLP per contiguous chunk along dim -2."""

    def __init__(self, mlps: Sequence[nn.Module], chunk_sizes: Sequence[int]):
        super().__init__()
        self.mlps = nn.ModuleList(mlps)
        self.chunk_sizes = list(chunk_sizes)

    def forward(self, x):
        chunks = torch.split(x, self.chunk_sizes, dim=-2)
        return torch.cat([m(c) for m, c in zip(self.mlps, chunks)], dim=-2)


class InteractionNet(nn.Module):
    """A.2.  edge_index (2,E): row 0 senders, row 1 receivers (any base)."""

    def __init__(
        self,
        edge_index: torch.Tensor,
        input_dim: int,
        update_edges: bool = True,
        hidden_layers: int = 1,
        hidden_dim: Optional[int] = None,
        edge_chunk_sizes: Optional[Sequence[int]] = None,
        aggr_chunk_sizes: Optional[Sequence[int]] = None,
        aggr: str = "sum",
        num_rec: Optional[int] = None,
    ):
        super().__init__()
        assert aggr in ("sum", "mean")
        hidden_dim = hidden_dim or input_dim
        ei = edge_index.clone()
        if num_rec is None:
            # upstream behaviour (A.2): re-base each row by its minimum, receivers counted from
            # the largest receiver id.  NOTE: this silently shifts sender ids when the
            # lowest-numbered sender has no edge — e.g. g2m, where the grid corner nodes lie at
            # dm/sqrt(2) > 0.67 dm from the nearest mesh node and send nothing.  The models below
            # therefore pass exact 0-based ids + num_rec (the intended semantics; mfai parity is
            # unpinned either way, see the header).
            ei = ei - ei.min(dim=1, keepdim=True)[0]
            self.num_rec = int(ei[1].max()) + 1
        else:
            self.num_rec = int(num_rec)
        ei[0] = ei[0] + self.num_rec  # senders after receivers in cat(rec, send)
        self.register_buffer("edge_index", ei, persistent=False)
        edge_recipe = [3 * input_dim] + [hidden_dim] * (hidden_layers + 1)
        aggr_recipe = [2 * input_dim] + [hidden_dim] * (hidden_layers + 1)
        if edge_chunk_sizes is None:
            self.edge_mlp = make_mlp(edge_recipe)
        else:
            self.edge_mlp = SplitMLPs(
                [make_mlp(edge_recipe) for _ in edge_chunk_sizes], edge_chunk_sizes
            )
        if aggr_chunk_sizes is None:
            self.aggr_mlp = make_mlp(aggr_recipe)
        else:
            self.aggr_mlp = SplitMLPs(
                [make_mlp(aggr_recipe) for _ in aggr_chunk_sizes], aggr_chunk_sizes
            )
        self.aggr = aggr
        self.update_edges = update_edges

    def forward(self, send_rep, rec_rep, edge_rep):
        node = torch.cat((rec_rep, send_rep), dim=-2)
        x_j = node.index_select(-2, self.edge_index[0])  # sender
        x_i = node.index_select(-2, self.edge_index[1])  # receiver
        msg = self.edge_mlp(torch.cat((edge_rep, x_j, x_i), dim=-1))
        aggr = torch.zeros(
            *msg.shape[:-2], self.num_rec, msg.shape[-1], dtype=msg.dtype, device=msg.device
        ).index_add_(-2, self.edge_index[1], msg)
        if self.aggr == "mean":
            deg = torch.bincount(self.edge_index[1], minlength=self.num_rec).clamp(min=1)
            aggr = aggr / deg.to(aggr.dtype).unsqueeze(-1)
        rec_rep = rec_rep + self.aggr_mlp(torch.cat((rec_rep, aggr), dim=-1))
        if self.update_edges:
            return rec_rep, edge_rep + msg
        return rec_rep


def _expand(x, B):
    return x.unsqueeze(0).expand(B, *x.shape)


class _BaseGraph(nn.Module):
    """A.3.  `graph` is a py4cast_b200.graph_build.MeshGraph (plain tensors)."""

    def __init__(self, in_channels, out_channels, graph, hidden_dims=64, hidden_layers=1):
        super().__init__()
        d = hidden_dims
        self.d, self.hl = d, hidden_layers
        end = [d] * (hidden_layers + 1)
        self.end = end
        self.register_buffer("g2m_features", graph.g2m.features.clone(), persistent=False)
        self.register_buffer("m2g_features", graph.m2g.features.clone(), persistent=False)
        self.grid_embedder = make_mlp([in_channels] + end)
        self.g2m_embedder = make_mlp([3] + end)
        self.m2g_embedder = make_mlp([3] + end)
        self.g2m_gnn = InteractionNet(
            graph.g2m.edge_index, d, hidden_layers=hidden_layers, update_edges=False,
            num_rec=graph.g2m.num_rec,
        )
        self.encoding_grid_mlp = make_mlp([d] + end)
        self.m2g_gnn = InteractionNet(
            graph.m2g.edge_index, d, hidden_layers=hidden_layers, update_edges=False,
            num_rec=graph.m2g.num_rec,
        )
        self.output_map = make_mlp(end + [out_channels], layer_norm=False)

    def forward(self, x):
        B = x.shape[0]
        grid_emb = self.grid_embedder(x)
        g2m_emb = self.g2m_embedder(self.g2m_features)
        m2g_emb = self.m2g_embedder(self.m2g_features)
        mesh_emb = self.embedd_mesh_nodes()
        mesh_rep = self.g2m_gnn(grid_emb, _expand(mesh_emb, B), _expand(g2m_emb, B))
        grid_rep = grid_emb + self.encoding_grid_mlp(grid_emb)
        mesh_rep = self.process_step(mesh_rep)
        grid_rep = self.m2g_gnn(mesh_rep, grid_rep, _expand(m2g_emb, B))
        return self.output_map(grid_rep)


class GraphLAM(_BaseGraph):
    """A.4."""

    def __init__(
        self, in_channels, out_channels, graph, hidden_dims=64, hidden_layers=1,
        processor_layers=4, mesh_aggr="sum",
    ):
        super().__init__(in_channels, out_channels, graph, hidden_dims, hidden_layers)
        assert not graph.hierarchical
        d = hidden_dims
        self.register_buffer(
            "mesh_static_features", graph.mesh_static_features[0].clone(), persistent=False
        )
        self.register_buffer("m2m_features", graph.m2m[0].features.clone(), persistent=False)
        self.mesh_embedder = make_mlp([2] + self.end)
        self.m2m_embedder = make_mlp([3] + self.end)
        self.processor = nn.ModuleList(
            [
                InteractionNet(
                    graph.m2m[0].edge_index, d, hidden_layers=hidden_layers, aggr=mesh_aggr,
                    num_rec=graph.m2m[0].num_rec,
                )
                for _ in range(processor_layers)
            ]
        )

    def embedd_mesh_nodes(self):
        return self.mesh_embedder(self.mesh_static_features)

    def process_step(self, mesh_rep):
        B = mesh_rep.shape[0]
        edge_rep = _expand(self.m2m_embedder(self.m2m_features), B)
        for net in self.processor:
            mesh_rep, edge_rep = net(mesh_rep, mesh_rep, edge_rep)
        return mesh_rep


class _BaseHiGraph(_BaseGraph):
    """A.5."""

    def __init__(self, in_channels, out_channels, graph, hidden_dims=64, hidden_layers=1):
        super().__init__(in_channels, out_channels, graph, hidden_dims, hidden_layers)
        assert graph.hierarchical
        d, hl = hidden_dims, hidden_layers
        L = graph.num_levels
        self.L = L
        self.level_mesh_sizes = [int(f.shape[0]) for f in graph.mesh_static_features]
        for l in range(L):
            self.register_buffer(f"mesh_static_{l}", graph.mesh_static_features[l].clone(), persistent=False)
            self.register_buffer(f"m2m_features_{l}", graph.m2m[l].features.clone(), persistent=False)
        for l in range(L - 1):
            self.register_buffer(f"up_features_{l}", graph.mesh_up[l].features.clone(), persistent=False)
            self.register_buffer(f"down_features_{l}", graph.mesh_down[l].features.clone(), persistent=False)
        self.mesh_embedders = nn.ModuleList([make_mlp([2] + self.end) for _ in range(L)])
        self.mesh_same_embedders = nn.ModuleList([make_mlp([3] + self.end) for _ in range(L)])
        self.mesh_up_embedders = nn.ModuleList([make_mlp([3] + self.end) for _ in range(L - 1)])
        self.mesh_down_embedders = nn.ModuleList([make_mlp([3] + self.end) for _ in range(L - 1)])
        self.mesh_init_gnns = nn.ModuleList(
            [InteractionNet(e.edge_index, d, hidden_layers=hl, num_rec=e.num_rec) for e in graph.mesh_up]
        )
        self.mesh_read_gnns = nn.ModuleList(
            [
                InteractionNet(e.edge_index, d, hidden_layers=hl, update_edges=False, num_rec=e.num_rec)
                for e in graph.mesh_down
            ]
        )

    def embedd_mesh_nodes(self):
        return self.mesh_embedders[0](self.mesh_static_0)

    def process_step(self, mesh_rep):
        B, L = mesh_rep.shape[0], self.L
        rep = [mesh_rep] + [
            _expand(self.mesh_embedders[l](getattr(self, f"mesh_static_{l}")), B)
            for l in range(1, L)
        ]
        same = [
            _expand(self.mesh_same_embedders[l](getattr(self, f"m2m_features_{l}")), B)
            for l in range(L)
        ]
        up = [
            _expand(self.mesh_up_embedders[l](getattr(self, f"up_features_{l}")), B)
            for l in range(L - 1)
        ]
        down = [
            _expand(self.mesh_down_embedders[l](getattr(self, f"down_features_{l}")), B)
            for l in range(L - 1)
        ]
        for l in range(1, L):  # init sweep (upward)
            rep[l], up[l - 1] = self.mesh_init_gnns[l - 1](rep[l - 1], rep[l], up[l - 1])
        rep, same, up, down = self.hi_processor_step(rep, same, up, down)
        cur = rep[-1]  # read-out sweep (downward)
        for l in range(L - 2, -1, -1):
            cur = self.mesh_read_gnns[l](cur, rep[l], down[l])
        return cur


class HiLAM(_BaseHiGraph):
    """A.6."""

    def __init__(
        self, in_channels, out_channels, graph, hidden_dims=64, hidden_layers=1,
        processor_layers=4, mesh_aggr="sum",
    ):
        super().__init__(in_channels, out_channels, graph, hidden_dims, hidden_layers)
        d, hl = hidden_dims, hidden_layers

        def nets(edge_sets):
            return nn.ModuleList(
                [InteractionNet(e.edge_index, d, hidden_layers=hl, num_rec=e.num_rec) for e in edge_sets]
            )

        self.mesh_down_gnns = nn.ModuleList([nets(graph.mesh_down) for _ in range(processor_layers)])
        self.mesh_down_same_gnns = nn.ModuleList([nets(graph.m2m) for _ in range(processor_layers)])
        self.mesh_up_gnns = nn.ModuleList([nets(graph.mesh_up) for _ in range(processor_layers)])
        self.mesh_up_same_gnns = nn.ModuleList([nets(graph.m2m) for _ in range(processor_layers)])

    def hi_processor_step(self, rep, same, up, down):
        L = self.L
        for dg, dsg, ug, usg in zip(
            self.mesh_down_gnns, self.mesh_down_same_gnns, self.mesh_up_gnns, self.mesh_up_same_gnns
        ):
            # down step
            rep[L - 1], same[L - 1] = dsg[L - 1](rep[L - 1], rep[L - 1], same[L - 1])
            for l in range(L - 2, -1, -1):
                new, down[l] = dg[l](rep[l + 1], rep[l], down[l])
                rep[l], same[l] = dsg[l](new, new, same[l])
            # up step
            rep[0], same[0] = usg[0](rep[0], rep[0], same[0])
            for l in range(1, L):
                new, up[l - 1] = ug[l - 1](rep[l - 1], rep[l], up[l - 1])
                rep[l], same[l] = usg[l](new, new, same[l])
        return rep, same, up, down


class HiLAMParallel(_BaseHiGraph):
    """A.7."""

    def __init__(
        self, in_channels, out_channels, graph, hidden_dims=64, hidden_layers=1,
        processor_layers=4, mesh_aggr="sum",
    ):
        super().__init__(in_channels, out_channels, graph, hidden_dims, hidden_layers)
        d, hl, L = hidden_dims, hidden_layers, self.L
        offs = [0]
        for s in self.level_mesh_sizes:
            offs.append(offs[-1] + s)
        parts = []
        for l in range(L):
            parts.append(graph.m2m[l].edge_index + offs[l])
        for l in range(L - 1):  # up: level l -> l+1
            ei = graph.mesh_up[l].edge_index
            parts.append(torch.stack([ei[0] + offs[l], ei[1] + offs[l + 1]]))
        for l in range(L - 1):  # down: level l+1 -> l
            ei = graph.mesh_down[l].edge_index
            parts.append(torch.stack([ei[0] + offs[l + 1], ei[1] + offs[l]]))
        self.edge_split_sections = [int(p.shape[1]) for p in parts]
        total = torch.cat(parts, dim=1)
        self.processor_nets = nn.ModuleList(
            [
                InteractionNet(
                    total, d, hidden_layers=hl,
                    edge_chunk_sizes=self.edge_split_sections,
                    aggr_chunk_sizes=self.level_mesh_sizes, num_rec=offs[-1],
                )
                for _ in range(processor_layers)
            ]
        )

    def hi_processor_step(self, rep, same, up, down):
        L = self.L
        if len(self.processor_nets) == 0:
            return rep, same, up, down
        node = torch.cat(rep, dim=1)
        edge = torch.cat(same + up + down, dim=1)
        for net in self.processor_nets:
            node, edge = net(node, node, edge)
        rep = list(torch.split(node, self.level_mesh_sizes, dim=1))
        secs = torch.split(edge, self.edge_split_sections, dim=1)
        return rep, list(secs[:L]), list(secs[L : 2 * L - 1]), list(secs[2 * L - 1 :])
