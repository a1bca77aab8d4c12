"""
GraphLAM / HiLAM / HiLAMParallel on the B200 kernels.

Same registry names, constructor contract, settings fields and forward
signature `(B, N_grid, F_in) -> (B, N_grid, F_out)` as the graph models the
reference trains through `mfai.pytorch.models` (reference call sites:
py4cast/lightning.py:265-279 construction, :591-596 forward;
tests/test_models.py:63-142 kwargs / rank_zero_setup; settings fields from
config/CLI/model/hilam.yaml:18-25).  The algorithm is the neural-lam one
(SURVEY.md §3.2, Appendix A); every MLP / message-passing step runs in
libpy4cast_b200 (py4cast_b200/ops.py), there is no eager fallback.

State-dict keys follow the nn.Sequential layout of `make_mlp`
(`<mlp>.0.weight`, `<mlp>.0.bias`, `<mlp>.2.weight`, `<mlp>.2.bias`, `<mlp>.3.weight`, `<mlp>.3.bias`).
"""

from __future__ import annotations

import os
import tempfile
from dataclasses import asdict, dataclass, fields
from pathlib import Path
from typing import List, Optional, Sequence, Tuple, Union

import torch
from torch import nn

from . import ops
from ._lib import HIDDEN
from .base import ModelABC, ModelType
from .graph_build import EdgeSet, MeshGraph, build_mesh_graph


# ------------------------------------------------------------------- settings
def _json_dataclass(cls):
    """Minimal stand-in for dataclasses_json.dataclass_json (absent here): to_dict/from_dict."""

    def to_dict(self):
        return {k: (str(v) if isinstance(v, Path) else v) for k, v in asdict(self).items()}

    @classmethod
    def from_dict(kls, d):
        return kls(**d)

    def to_json(self):
        import json

        return json.dumps(self.to_dict())

    @classmethod
    def from_json(kls, s):
        import json

        return kls(**json.loads(s))

    cls.to_dict, cls.from_dict, cls.to_json, cls.from_json = to_dict, from_dict, to_json, from_json
    return cls


@_json_dataclass
@dataclass(slots=True)
class GraphLamSettings:
    """Settings for graph-based models (fields of config/CLI/model/hilam.yaml:18-25)."""

    tmp_dir: Union[Path, str] = Path(tempfile.gettempdir())
    hidden_dims: int = 64
    hidden_layers: int = 1
    processor_layers: int = 4
    mesh_aggr: str = "sum"
    use_checkpointing: bool = False
    offload_to_cpu: bool = False
    # cap on the number of mesh levels (neural-lam `--levels`); None = all levels
    n_max_levels: Optional[int] = None


# ------------------------------------------------------------------ MLP module
class FusedMLP(nn.Sequential):
    """`make_mlp([in, d, out])`: Linear, SiLU, Linear, [LayerNorm] — executed as ONE kernel."""

    def __init__(self, in_dim: int, hidden: int, out_dim: int, layer_norm: bool = True,
                 widths: Optional[Sequence[int]] = None):
        layers: List[nn.Module] = [nn.Linear(in_dim, hidden), nn.SiLU(), nn.Linear(hidden, out_dim)]
        if layer_norm:
            layers.append(nn.LayerNorm(out_dim))
        super().__init__(*layers)
        self.has_ln = layer_norm
        # widths of the concatenated inputs of layer 1 (edge MLP: edge, sender, receiver)
        self.widths = tuple(widths) if widths is not None else (in_dim,)
        assert sum(self.widths) == in_dim
        self._meta: Optional[ops.MlpMeta] = None

    @property
    def meta(self) -> ops.MlpMeta:
        """Packed-weight image / deferred weight-gradient state of this MLP (tensor-core engine)."""
        if self._meta is None:
            self._meta = ops.MlpMeta(self.weights(), self.widths, self[2].out_features)
        return self._meta

    def weights(self) -> Tuple[Optional[torch.Tensor], ...]:
        ln = self[3] if self.has_ln else None
        return (
            self[0].weight, self[0].bias, self[2].weight, self[2].bias,
            ln.weight if ln is not None else None, ln.bias if ln is not None else None,
        )

    def forward(self, x: torch.Tensor, residual: bool = False, share_grads: bool = False) -> torch.Tensor:
        return ops.fused_mlp(x, self.weights(), residual=residual, meta=self.meta, share_grads=share_grads)


def make_mlp(blueprint: Sequence[int], layer_norm: bool = True, widths: Optional[Sequence[int]] = None) -> FusedMLP:
    if len(blueprint) != 3:
        raise NotImplementedError(
            f"fused MLP kernels implement hidden_layers=1 (blueprint of 3 widths), got {list(blueprint)}"
        )
    if blueprint[1] != HIDDEN:
        raise NotImplementedError(f"fused MLP kernels are built for hidden_dims={HIDDEN}, got {blueprint[1]}")
    return FusedMLP(blueprint[0], blueprint[1], blueprint[2], layer_norm, widths)


class SplitMLPs(nn.Module):
    """One MLP per contiguous row chunk (HiLAMParallel); only a weight container here."""

    def __init__(self, mlps: Sequence[FusedMLP], chunk_sizes: Sequence[int]):
        super().__init__()
        self.mlps = nn.ModuleList(mlps)
        self.chunk_sizes = list(chunk_sizes)


class InteractionNet(nn.Module):
    """One message-passing step; see ops._InteractionNetFn.  edge_index: (2,E) local 0-based ids."""

    def __init__(
        self,
        edge_set: EdgeSet,
        input_dim: int,
        update_edges: bool = True,
        hidden_layers: int = 1,
        hidden_dim: Optional[int] = None,
        edge_chunk_sizes: Optional[Sequence[int]] = None,
        aggr_chunk_sizes: Optional[Sequence[int]] = None,
        aggr: str = "sum",
    ):
        super().__init__()
        if aggr not in ("sum", "mean"):
            raise ValueError(f"Unknown aggregation method: {aggr}")
        hidden_dim = hidden_dim or input_dim
        if hidden_layers != 1:
            raise NotImplementedError("fused InteractionNet kernels implement hidden_layers=1")
        self.register_buffer("edge_index", edge_set.edge_index.clone(), persistent=False)
        self.num_send, self.num_rec = edge_set.num_send, edge_set.num_rec
        edge_bp = [3 * input_dim, hidden_dim, hidden_dim]
        aggr_bp = [2 * input_dim, hidden_dim, hidden_dim]
        ew, aw = [input_dim] * 3, [input_dim] * 2
        if edge_chunk_sizes is None:
            self.edge_mlp = make_mlp(edge_bp, widths=ew)
        else:
            self.edge_mlp = SplitMLPs([make_mlp(edge_bp, widths=ew) for _ in edge_chunk_sizes], edge_chunk_sizes)
        if aggr_chunk_sizes is None:
            self.aggr_mlp = make_mlp(aggr_bp, widths=aw)
        else:
            self.aggr_mlp = SplitMLPs([make_mlp(aggr_bp, widths=aw) for _ in aggr_chunk_sizes], aggr_chunk_sizes)
        self.edge_chunk_sizes = None if edge_chunk_sizes is None else list(edge_chunk_sizes)
        self.aggr_chunk_sizes = None if aggr_chunk_sizes is None else list(aggr_chunk_sizes)
        self.aggr = aggr
        self.update_edges = update_edges
        self._gi = None

    def graph_index(self) -> ops.GraphIndex:
        if self._gi is None or self._gi.send_idx.device != self.edge_index.device:
            if not self.edge_index.is_cuda:
                raise ops._lib.P4CError(
                    "py4cast_b200 graph models run on CUDA only (no CPU fallback): move the "
                    "model to a cuda device"
                )
            self._gi = ops.GraphIndex.build(
                self.edge_index, self.num_send, self.num_rec,
                self.edge_chunk_sizes, self.aggr_chunk_sizes,
            )
        return self._gi

    def _weights(self) -> List[torch.Tensor]:
        ws: List[torch.Tensor] = []
        for m in (self.edge_mlp, self.aggr_mlp):
            for mlp in (m.mlps if isinstance(m, SplitMLPs) else [m]):
                ws.extend(mlp.weights())
        return ws

    def _metas(self) -> List[ops.MlpMeta]:
        ms: List[ops.MlpMeta] = []
        for m in (self.edge_mlp, self.aggr_mlp):
            for mlp in (m.mlps if isinstance(m, SplitMLPs) else [m]):
                ms.append(mlp.meta)
        return ms

    def forward(self, send_rep, rec_rep, edge_rep, share_grads: bool = False):
        return ops.interaction_net(
            self.graph_index(), send_rep, rec_rep, edge_rep, self._weights(),
            self.update_edges, self.aggr == "mean", self._metas(), share_grads,
        )


# ---------------------------------------------------------------- graph storage
def _graph_path(tmp_dir, hierarchical: bool, grid_shape, n_max_levels) -> Path:
    tag = "hi" if hierarchical else "flat"
    return Path(tmp_dir) / f"p4c_b200_graph_{tag}_{grid_shape[0]}x{grid_shape[1]}_L{n_max_levels or 0}.pt"


def _meshgrid_digest(meshgrid: torch.Tensor) -> str:
    """Identifies the domain a cached graph was built for (two domains of the same shape but other
    coordinates must not share a graph)."""
    import hashlib

    t = torch.as_tensor(meshgrid).detach().to("cpu", torch.float64).contiguous()
    return hashlib.sha256(t.numpy().tobytes()).hexdigest()[:16]


def _graph_to_dict(g: MeshGraph) -> dict:
    def es(e: EdgeSet):
        return {"edge_index": e.edge_index, "features": e.features, "num_send": e.num_send, "num_rec": e.num_rec}

    return {
        "hierarchical": g.hierarchical, "grid_shape": tuple(g.grid_shape), "level_sizes": list(g.level_sizes),
        "mesh_static_features": g.mesh_static_features, "m2m": [es(e) for e in g.m2m],
        "g2m": es(g.g2m), "m2g": es(g.m2g), "mesh_up": [es(e) for e in g.mesh_up],
        "mesh_down": [es(e) for e in g.mesh_down],
    }


def _graph_from_dict(d: dict) -> MeshGraph:
    def es(x):
        return EdgeSet(x["edge_index"], x["features"], x["num_send"], x["num_rec"])

    return MeshGraph(
        hierarchical=d["hierarchical"], grid_shape=tuple(d["grid_shape"]), level_sizes=list(d["level_sizes"]),
        mesh_static_features=d["mesh_static_features"], m2m=[es(e) for e in d["m2m"]], g2m=es(d["g2m"]),
        m2g=es(d["m2g"]), mesh_up=[es(e) for e in d["mesh_up"]], mesh_down=[es(e) for e in d["mesh_down"]],
    )


# ------------------------------------------------------------------ base model
class BaseGraphModel(ModelABC, nn.Module):
    settings_kls = GraphLamSettings
    onnx_supported: bool = False
    supported_num_spatial_dims = (1,)
    num_spatial_dims: int = 1
    features_last: bool = True
    model_type = ModelType.GRAPH
    register: bool = False
    hierarchical: bool = False

    @classmethod
    def rank_zero_setup(cls, settings: GraphLamSettings, meshgrid: torch.Tensor) -> None:
        """Build the mesh graph for `meshgrid` (2,H,W) and store it under settings.tmp_dir
        (reference lightning.py:141-144: called once, on rank 0, before construction).
        The write is atomic so concurrent ranks may safely call it too."""
        graph = build_mesh_graph(meshgrid, cls.hierarchical, settings.n_max_levels)
        path = _graph_path(settings.tmp_dir, cls.hierarchical, graph.grid_shape, settings.n_max_levels)
        path.parent.mkdir(parents=True, exist_ok=True)
        fd, tmp = tempfile.mkstemp(dir=str(path.parent), suffix=".tmp")
        os.close(fd)
        payload = _graph_to_dict(graph)
        payload["meshgrid_digest"] = _meshgrid_digest(meshgrid)
        torch.save(payload, tmp)
        os.replace(tmp, path)

    def __init__(
        self,
        in_channels: int,
        out_channels: int,
        input_shape: Tuple[int, ...],
        settings: GraphLamSettings = None,
        *args,
        graph: Optional[MeshGraph] = None,
        **kwargs,
    ):
        super().__init__()
        settings = settings if settings is not None else GraphLamSettings()
        self._settings = settings
        self.in_channels, self.out_channels, self.input_shape = in_channels, out_channels, tuple(input_shape)
        if settings.use_checkpointing or settings.offload_to_cpu:
            # memory knobs of the reference (activation recomputation / host offload for 32 GB GPUs): they change
            # where activations live, never a result.  Configurations that set them keep running here; the fused
            # backward already recomputes the layer-1 activations and LayerNorm statistics of every large call, and
            # what is saved (about 0.35 GB per sample and AR step at the Titan size) stays in HBM.
            import warnings

            warnings.warn(
                "use_checkpointing / offload_to_cpu have no effect in py4cast_b200 (same results; the fused backward "
                "recomputes activations already and the saved ones stay in the 180 GB of HBM)", stacklevel=2)
        if graph is None:
            path = _graph_path(settings.tmp_dir, self.hierarchical, self.input_shape, settings.n_max_levels)
            if not path.exists():
                raise FileNotFoundError(
                    f"{path} not found: call {type(self).__name__}.rank_zero_setup(settings, meshgrid) first"
                )
            # tensors, lists, ints and strings only: no pickle execution from a shared temp dir
            payload = torch.load(path, map_location="cpu", weights_only=True)
            self.graph_digest = payload.get("meshgrid_digest")
            graph = _graph_from_dict(payload)
        if graph.hierarchical != self.hierarchical:
            raise ValueError("graph hierarchy flag does not match the model class")
        if graph.num_grid_nodes != self.input_shape[0] * self.input_shape[1]:
            raise ValueError(f"graph built for grid {graph.grid_shape}, model input_shape {self.input_shape}")
        self.graph_meta = {"level_sizes": graph.level_sizes, "grid_shape": graph.grid_shape}

        d, hl = settings.hidden_dims, settings.hidden_layers
        self.d = d
        end = [d] * (hl + 1)
        self.mlp_blueprint_end = end
        self._feat_cache: dict = {}
        self._reg_feats("g2m_features", graph.g2m.features)
        self._reg_feats("m2g_features", graph.m2g.features)
        self.grid_embedder = make_mlp([in_channels] + end)
        self.g2m_embedder = make_mlp([3] + end)
        self.m2g_embedder = make_mlp([3] + end)
        self.g2m_gnn = InteractionNet(graph.g2m, d, hidden_layers=hl, update_edges=False)
        self.encoding_grid_mlp = make_mlp([d] + end)
        self.m2g_gnn = InteractionNet(graph.m2g, d, hidden_layers=hl, update_edges=False)
        self.output_map = make_mlp(end + [out_channels], layer_norm=False)
        self._build_mesh(graph, settings)
        self.check_required_attributes()

    # static (non-trainable) features are kept in fp32 buffers + per-dtype casted copies
    def _reg_feats(self, name: str, t: torch.Tensor) -> None:
        self.register_buffer(name, t.clone().contiguous(), persistent=False)

    def _feats(self, name: str, dtype: torch.dtype) -> torch.Tensor:
        base = getattr(self, name)
        if dtype == torch.float32:
            return base
        key = (name, dtype, base.device)
        if key not in self._feat_cache:
            # rows padded with zeros to 16 bytes: the tensor-core MLP fetches them as one chunk
            padded = ops.empty_rows(base.shape, dtype, base.device, zero=True)
            padded.copy_(ops.cast(base, dtype))
            self._feat_cache[key] = padded
        return self._feat_cache[key]

    @property
    def settings(self) -> GraphLamSettings:
        return self._settings

    # ---- rollout-level cache of the batch-independent embeddings ---------------------------
    # Edge / mesh-node embeddings depend on the weights and on static graph features only.  The
    # reference recomputes them in every forward call (T times per AR rollout, 1.3 M edges for
    # m2g); between `begin_rollout()` and `end_rollout()` they are computed once and shared by
    # all AR steps (autograd accumulates their gradient over the steps).
    def pack_weights(self) -> None:
        """bf16 weight images of every MLP of the model, ONE kernel launch (the tensor-core kernels
        then stage weights with 16-byte copies).  Must run after every weight update and before the
        next forward: `forward` does it itself outside a rollout, `begin_rollout` once per rollout."""
        packer = getattr(self, "_packer", None)
        if packer is None:
            packer = ops.WeightPacker([m.meta for m in self.modules() if isinstance(m, FusedMLP)])
            self._packer = packer
        packer.pack()

    def begin_rollout(self, dtype: Optional[torch.dtype] = None) -> None:
        """Re-entrant: an outer begin_rollout (the trainer, around several concurrent micro-batch
        rollouts) keeps one cache / one weight packing for all inner ones."""
        depth = getattr(self, "_rollout_depth", 0)
        self._rollout_depth = depth + 1
        if depth > 0:
            return
        self._rollout_cache = {}
        if dtype is None or dtype == torch.bfloat16:
            self.pack_weights()

    def end_rollout(self) -> None:
        self._rollout_depth = max(0, getattr(self, "_rollout_depth", 1) - 1)
        if self._rollout_depth == 0:
            self._rollout_cache = None

    def precompute_static(self, dt: torch.dtype) -> None:
        """Fills the rollout cache with every batch-independent embedding (on the current stream),
        so that rollouts running on other streams only read them."""
        if getattr(self, "_rollout_cache", None) is None:
            raise RuntimeError("precompute_static() needs begin_rollout() first")
        self._static_embedding("g2m", self.g2m_embedder, "g2m_features", dt)
        self._static_embedding("m2g", self.m2g_embedder, "m2g_features", dt)
        self.embedd_mesh_nodes(dt)
        self._static_mesh_embeddings(dt)

    def _static_mesh_embeddings(self, dt: torch.dtype) -> None:
        pass

    def _static_embedding(self, key: str, mlp: "FusedMLP", feats_name: str, dt: torch.dtype) -> torch.Tensor:
        cache = getattr(self, "_rollout_cache", None)
        if cache is None:
            return mlp(self._feats(feats_name, dt))
        k = (key, dt)
        if k not in cache:
            cache[k] = mlp(self._feats(feats_name, dt))
        return cache[k]

    def _compute_dtype(self, x: torch.Tensor) -> torch.dtype:
        # reference: the model runs under torch.amp.autocast("cuda", dtype=str_to_dtype[precision])
        # (lightning.py:485-487); bf16 autocast or a bf16 input selects bf16 activation storage.
        if torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16:
            return torch.bfloat16
        if x.dtype == torch.bfloat16:
            return torch.bfloat16
        return torch.float32

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() != 3:
            raise ValueError(f"graph models expect (B, N_grid, F_in), got {tuple(x.shape)}")
        dt = self._compute_dtype(x)
        if getattr(self, "_rollout_cache", None) is None and dt == torch.bfloat16:
            self.pack_weights()
        x = ops.cast_autograd(x if x.stride(-1) == 1 else x.contiguous(), dt)
        grid_emb = self.grid_embedder(x)
        g2m_emb = self._static_embedding("g2m", self.g2m_embedder, "g2m_features", dt)
        m2g_emb = self._static_embedding("m2g", self.m2g_embedder, "m2g_features", dt)
        mesh_emb = self.embedd_mesh_nodes(dt)
        # grid_emb has exactly two consumers, the g2m InteractionNet (senders) and the grid encoder: their input
        # gradients are summed by g2m's by-sender kernel (ops: gradient hand-over), not by an element-wise pass
        mesh_rep = self.g2m_gnn(grid_emb, mesh_emb, g2m_emb, share_grads=True)
        grid_rep = self.encoding_grid_mlp(grid_emb, residual=True, share_grads=True)
        mesh_rep = self.process_step(mesh_rep, dt)
        grid_rep = self.m2g_gnn(mesh_rep, grid_rep, m2g_emb)
        return self.output_map(grid_rep)


class GraphLAM(BaseGraphModel):
    """Flat multi-scale mesh."""

    register = True
    hierarchical = False

    def _build_mesh(self, graph: MeshGraph, settings: GraphLamSettings) -> None:
        d, hl = self.d, settings.hidden_layers
        self._reg_feats("mesh_static_features", graph.mesh_static_features[0])
        self._reg_feats("m2m_features", graph.m2m[0].features)
        self.mesh_embedder = make_mlp([2] + self.mlp_blueprint_end)
        self.m2m_embedder = make_mlp([3] + self.mlp_blueprint_end)
        self.processor = nn.ModuleList(
            [
                InteractionNet(graph.m2m[0], d, hidden_layers=hl, aggr=settings.mesh_aggr)
                for _ in range(settings.processor_layers)
            ]
        )

    def embedd_mesh_nodes(self, dt):
        return self._static_embedding("mesh", self.mesh_embedder, "mesh_static_features", dt)

    def _static_mesh_embeddings(self, dt):
        self._static_embedding("m2m", self.m2m_embedder, "m2m_features", dt)

    def process_step(self, mesh_rep, dt):
        edge_rep = self._static_embedding("m2m", self.m2m_embedder, "m2m_features", dt)
        for net in self.processor:
            mesh_rep, edge_rep = net(mesh_rep, mesh_rep, edge_rep)
        return mesh_rep


class BaseHiGraphModel(BaseGraphModel):
    """Hierarchical mesh."""

    hierarchical = True

    def _build_mesh(self, graph: MeshGraph, settings: GraphLamSettings) -> None:
        d, hl = self.d, settings.hidden_layers
        L = graph.num_levels
        self.num_levels = L
        self.level_mesh_sizes = [int(f.shape[0]) for f in graph.mesh_static_features]
        for l in range(L):
            self._reg_feats(f"mesh_static_{l}", graph.mesh_static_features[l])
            self._reg_feats(f"m2m_features_{l}", graph.m2m[l].features)
        for l in range(L - 1):
            self._reg_feats(f"up_features_{l}", graph.mesh_up[l].features)
            self._reg_feats(f"down_features_{l}", graph.mesh_down[l].features)
        end = self.mlp_blueprint_end
        self.mesh_embedders = nn.ModuleList([make_mlp([2] + end) for _ in range(L)])
        self.mesh_same_embedders = nn.ModuleList([make_mlp([3] + end) for _ in range(L)])
        self.mesh_up_embedders = nn.ModuleList([make_mlp([3] + end) for _ in range(L - 1)])
        self.mesh_down_embedders = nn.ModuleList([make_mlp([3] + end) for _ in range(L - 1)])
        self.mesh_init_gnns = nn.ModuleList([InteractionNet(e, d, hidden_layers=hl) for e in graph.mesh_up])
        self.mesh_read_gnns = nn.ModuleList(
            [InteractionNet(e, d, hidden_layers=hl, update_edges=False) for e in graph.mesh_down]
        )
        self._build_processor(graph, settings)

    def embedd_mesh_nodes(self, dt):
        return self._static_embedding("mesh0", self.mesh_embedders[0], "mesh_static_0", dt)

    def _static_mesh_embeddings(self, dt):
        L = self.num_levels
        emb = self._static_embedding
        higher = [emb(f"mesh{l}", self.mesh_embedders[l], f"mesh_static_{l}", dt) for l in range(1, L)]
        same = [emb(f"same{l}", self.mesh_same_embedders[l], f"m2m_features_{l}", dt) for l in range(L)]
        up = [emb(f"up{l}", self.mesh_up_embedders[l], f"up_features_{l}", dt) for l in range(L - 1)]
        down = [emb(f"down{l}", self.mesh_down_embedders[l], f"down_features_{l}", dt) for l in range(L - 1)]
        return higher, same, up, down

    def process_step(self, mesh_rep, dt):
        L = self.num_levels
        higher, same, up, down = self._static_mesh_embeddings(dt)
        rep = [mesh_rep] + higher
        for l in range(1, L):  # init sweep, upward
            rep[l], up[l - 1] = self.mesh_init_gnns[l - 1](rep[l - 1], rep[l], up[l - 1], share_grads=True)
        rep, same, up, down = self.hi_processor_step(rep, same, up, down)
        cur = rep[-1]  # read-out sweep, downward
        for l in range(L - 2, -1, -1):
            cur = self.mesh_read_gnns[l](cur, rep[l], down[l], share_grads=True)
        return cur


class HiLAM(BaseHiGraphModel):
    """Sequential hierarchical processor."""

    register = True

    def _build_processor(self, graph: MeshGraph, settings: GraphLamSettings) -> None:
        d, hl, pl = self.d, settings.hidden_layers, settings.processor_layers

        def nets(edge_sets):
            return nn.ModuleList([InteractionNet(e, d, hidden_layers=hl) for e in edge_sets])

        self.mesh_down_gnns = nn.ModuleList([nets(graph.mesh_down) for _ in range(pl)])
        self.mesh_down_same_gnns = nn.ModuleList([nets(graph.m2m) for _ in range(pl)])
        self.mesh_up_gnns = nn.ModuleList([nets(graph.mesh_up) for _ in range(pl)])
        self.mesh_up_same_gnns = nn.ModuleList([nets(graph.m2m) for _ in range(pl)])

    def hi_processor_step(self, rep, same, up, down):
        L = self.num_levels
        for dg, dsg, ug, usg in zip(
            self.mesh_down_gnns, self.mesh_down_same_gnns, self.mesh_up_gnns, self.mesh_up_same_gnns
        ):
            rep[L - 1], same[L - 1] = dsg[L - 1](rep[L - 1], rep[L - 1], same[L - 1])
            for l in range(L - 2, -1, -1):
                new, down[l] = dg[l](rep[l + 1], rep[l], down[l], share_grads=True)
                rep[l], same[l] = dsg[l](new, new, same[l])
            rep[0], same[0] = usg[0](rep[0], rep[0], same[0])
            for l in range(1, L):
                new, up[l - 1] = ug[l - 1](rep[l - 1], rep[l], up[l - 1], share_grads=True)
                rep[l], same[l] = usg[l](new, new, same[l])
        return rep, same, up, down


class HiLAMParallel(BaseHiGraphModel):
    """All levels' edges in one message-passing call per layer."""

    register = True

    def _build_processor(self, graph: MeshGraph, settings: GraphLamSettings) -> None:
        d, hl, L = self.d, settings.hidden_layers, self.num_levels
        offs = [0]
        for s in self.level_mesh_sizes:
            offs.append(offs[-1] + s)
        parts = [graph.m2m[l].edge_index + offs[l] for l in range(L)]
        for l in range(L - 1):
            ei = graph.mesh_up[l].edge_index
            parts.append(torch.stack([ei[0] + offs[l], ei[1] + offs[l + 1]]))
        for l in range(L - 1):
            ei = graph.mesh_down[l].edge_index
            parts.append(torch.stack([ei[0] + offs[l + 1], ei[1] + offs[l]]))
        self.edge_split_sections = [int(p.shape[1]) for p in parts]
        total = EdgeSet(torch.cat(parts, dim=1), torch.zeros(0, 3), offs[-1], offs[-1])
        self.processor_nets = nn.ModuleList(
            [
                InteractionNet(
                    total, d, hidden_layers=hl,
                    edge_chunk_sizes=self.edge_split_sections, aggr_chunk_sizes=self.level_mesh_sizes,
                )
                for _ in range(settings.processor_layers)
            ]
        )

    @staticmethod
    def _batched(t: torch.Tensor, B: int) -> torch.Tensor:
        return t if t.dim() == 3 else t.unsqueeze(0).expand(B, *t.shape)

    def hi_processor_step(self, rep, same, up, down):
        L = self.num_levels
        if len(self.processor_nets) == 0:
            return rep, same, up, down
        B = rep[0].shape[0]
        node = torch.cat([self._batched(r, B) for r in rep], dim=1)
        edge = torch.cat([self._batched(e, B) for e in (same + up + down)], dim=1)
        for net in self.processor_nets:
            node, edge = net(node, node, edge)
        rep = [t.contiguous() for t in torch.split(node, self.level_mesh_sizes, dim=1)]
        secs = [t.contiguous() for t in torch.split(edge, self.edge_split_sections, dim=1)]
        return rep, secs[:L], secs[L: 2 * L - 1], secs[2 * L - 1:]


GRAPH_MODELS = {"GraphLAM": GraphLAM, "HiLAM": HiLAM, "HiLAMParallel": HiLAMParallel}
