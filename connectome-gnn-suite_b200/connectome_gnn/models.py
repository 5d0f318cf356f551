"""GCN and GraphSAGE connectome classifiers on the sm_100a kernels.

Host-side mirror of reference ``connectome_gnn/models.py``: same class names, constructor
arguments, attributes (``convs``, ``batch_norms``, ``classifier``, ``dropout``) and
``state_dict`` keys, so weights move freely between the two implementations.  The modules
only *hold parameters*; the arithmetic of ``encode``/``forward`` is two autograd nodes
(:class:`_EncodeFn`, :class:`_HeadFn`) that drive the C ABI:

====================  =========================================================================
reference lines       replaced by
====================  =========================================================================
models.py:84-114      ``cgnn_gcn_layer_fwd`` / ``_bwd`` (projection + normalised aggregation)
models.py:136-152     ``cgnn_sage_layer_fwd`` / ``_bwd`` (weighted-mean aggregation + projection)
models.py:208-210,    BatchNorm + ReLU + dropout applied on load by the *next* kernel;
  260-261             statistics come out of the producing kernel (``cgnn_bn_finalize``)
models.py:57-59,211   ``cgnn_pool_fwd``;  backward folded into the top layer's backward
models.py:196-201     ``cgnn_head_fwd`` / ``_bwd``
====================  =========================================================================

Under ``torch.distributed`` (one process per GPU, subjects sharded by the loader) the BatchNorm
statistics and their backward sums are combined across ranks so that the result equals the
reference's single-process batch semantics; parameter gradients are all-reduced by ``Trainer``.
"""

from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import _engine
from ._engine import Act, BnBwd
from .graph import ConnectomeBatch

__all__ = ["GCNLayer", "SAGELayer", "GCNConnectome", "GraphSAGEConnectome"]

FUSED_EVAL_MAX_ROWS = 2 * 148 * 384   # two units per SM: below this the per-layer path is bound by its ~12 launches


# ---------------------------------------------------------------------------
# distributed helpers (no-ops in a single process)
# ---------------------------------------------------------------------------

def _world(group=None) -> int:
    """Ranks that share BatchNorm statistics; ``group=False`` (``model.process_group = False``) = this process alone,
    whatever torch.distributed says (single-process semantics under a multi-process launcher)."""
    import torch.distributed as dist
    if group is False:
        return 1
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group)
    return 1


def _merge_stats_across_ranks(eng, stats: torch.Tensor, channels: int, group=None, peers=None, layer: int = 0) -> torch.Tensor:
    """SyncBN forward: gather every rank's {count, mean, M2} and merge them identically everywhere - over NVLink peer
    memory in one kernel when ``peers`` (a PeerExchange) is given, else NCCL all-gather + merge kernel."""
    world = _world(group)
    if world == 1:
        return stats
    if peers is not None:
        return peers.merge_stats(layer, stats, channels)
    import torch.distributed as dist
    flat = torch.empty(world * stats.numel(), dtype=stats.dtype, device=stats.device)   # 1-D: gloo insists on it
    dist.all_gather_into_tensor(flat, stats.contiguous().reshape(-1), group=group)
    return eng.bn_merge_stats(flat.view(world, stats.numel()), channels)


def _sum_across_ranks(sums: Optional[torch.Tensor], sums64: Optional[torch.Tensor], group=None, peers=None, layer: int = 0):
    """SyncBN backward: the [sum dy, sum dy x^] records of all ranks added up - the float64 record (the one the kernels
    read: BatchNorm's backward subtracts these means from every row, so the sums are carried unrounded from the reduction
    kernel to their consumer, also across ranks) and from it the fp32 copy that becomes d beta / d gamma."""
    if sums64 is not None and _world(group) > 1:
        if peers is not None:
            sums64 = peers.sum(layer, sums64)
        else:
            import torch.distributed as dist
            dist.all_reduce(sums64, group=group)
        sums.copy_(sums64)
    return sums, sums64


def _draw_seed(device=None) -> int:
    """Dropout stream seed.  On a CUDA device it is cut from that device's default CUDA generator exactly the way
    torch's own CUDA dropout consumes it: (seed, philox offset) is read on the host and the offset advanced - so
    ``torch.manual_seed`` governs the masks, successive forwards differ, every data-parallel rank seeded alike draws
    alike, and the global CPU generator (whose ``randperm`` stream defines batch membership, reference ``graph.py:193``)
    is never touched, as in the reference running on CUDA.  Without CUDA (test-only simulator) the CPU generator is used."""
    if device is not None and torch.device(device).type == "cuda" and torch.cuda.is_available():
        dev = torch.device(device)
        torch.cuda.init()          # default_generators is filled by the lazy initialisation
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        gen = torch.cuda.default_generators[idx]
        seed, off = int(gen.initial_seed()), int(gen.get_offset())
        gen.set_offset(off + 4)
        x = (seed * 0x9E3779B97F4A7C15 + (off // 4 + 1) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        x ^= x >> 31
        x = (x * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return (x ^ (x >> 29)) & 0x3FFFFFFFFFFFFFFF
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


# ---------------------------------------------------------------------------
# encode = L x (conv -> BN -> [ReLU] -> dropout) -> mean pool, as one autograd node
# ---------------------------------------------------------------------------

class _EncodeFn(torch.autograd.Function):
    """``emb = pool(stack(x))``.  ``cfg`` carries everything that is not a differentiable tensor."""

    @staticmethod
    def forward(ctx, cfg: dict, x: torch.Tensor, *params: torch.Tensor):
        eng = cfg["engine"]
        batch: ConnectomeBatch = cfg["batch"]
        kind, training, p = cfg["kind"], cfg["training"], cfg["dropout"]
        csr, ptr, B = batch.csr, batch.ptr, batch.num_graphs
        relu_after_bn = kind == "gcn"          # SAGE applies ReLU inside the layer (models.py:152)
        p_eff = p if training else 0.0
        seed = cfg["seed"]

        t, act = x.contiguous(), Act()
        saved = []
        L = len(cfg["bns"])
        for l, bn in enumerate(cfg["bns"]):
            W, b, gamma, beta = (q.contiguous() for q in params[4 * l: 4 * l + 4])
            if l == L - 1 and l > 0 and kind == "gcn" and not training and cfg.get("forward_only"):
                # eval mode, nobody will call backward: the last layer folds its BatchNorm (running statistics), ReLU and
                # the mean-pool readout into the layer kernel - z_L is never written
                scale, shift, mean, rstd = eng.bn_eval_affine(gamma, beta, bn.running_mean, bn.running_var, bn.eps)
                emb = eng.gcn_layer_fwd_pool(t, act, W, b, csr, ptr, B, Act(scale, shift, True, 0.0, 0, l, batch.row_base))
                if emb is not None:
                    ctx.cfg, ctx.saved, ctx.final_act, ctx.x_needs_grad = cfg, None, None, False
                    return emb
            z, stats, agg = eng.layer_fwd(kind, t, act, W, b, csr, ptr, B, want_stats=training)
            if training:
                stats = _merge_stats_across_ranks(eng, stats, W.shape[0], cfg["group"], cfg.get("peers"), l)
                if bn.momentum is None:    # torch: cumulative moving average, factor 1 / num_batches_tracked (after increment)
                    momentum = 1.0 / float(int(bn.num_batches_tracked) + 1)
                else:
                    momentum = bn.momentum
                scale, shift, mean, rstd = eng.bn_finalize(stats, gamma, beta, bn.eps, momentum, bn.running_mean,
                                                           bn.running_var, bn.num_batches_tracked)
            else:
                scale, shift, mean, rstd = eng.bn_eval_affine(gamma, beta, bn.running_mean, bn.running_var, bn.eps)
            saved.append((t, act, z, W, scale, mean, rstd, agg))
            t = z
            act = Act(scale, shift, relu_after_bn, p_eff, seed, l, batch.row_base, cfg.get("salt") if p_eff > 0.0 else None)
        emb = eng.pool_fwd(t, act, ptr, B)

        ctx.cfg, ctx.saved, ctx.final_act = cfg, saved, act
        ctx.x_needs_grad = x.requires_grad
        return emb

    @staticmethod
    def backward(ctx, demb: torch.Tensor):
        cfg, saved = ctx.cfg, ctx.saved
        eng = cfg["engine"]
        batch: ConnectomeBatch = cfg["batch"]
        kind, training, group = cfg["kind"], cfg["training"], cfg["group"]
        csr, ptr, B = batch.csr, batch.ptr, batch.num_graphs
        count = float(batch.global_num_nodes if batch.global_num_nodes is not None else batch.num_nodes)
        L = len(saved)
        demb = demb.contiguous()

        grads: list = [None] * (4 * L)
        # With a flat gradient buffer (Trainer.enable_fused_step) the kernels write every parameter gradient straight into its
        # slice - the BatchNorm pair [d beta; d gamma] is the [2, C] block of backward sums itself - and autograd receives
        # fresh views of those slices: no per-parameter copy, one all-reduce, one optimizer kernel.
        flat = cfg.get("flat")
        layer_params = cfg.get("layer_params")        # [(W, b, gamma, beta)] Parameter objects per layer (flat mode)
        bn_block = (lambda l: flat.bn_block(layer_params[l][3], layer_params[l][2])) if flat is not None else (lambda l: None)
        # BatchNorm backward sums of the top layer come from a standalone pass over z_L
        t_in, act_in, z, W, scale, mean, rstd, agg = saved[-1]
        act_out = ctx.final_act
        sums, sums64 = eng.bn_bwd_sums(z, act_out, mean, rstd, None, demb, ptr, B, out=bn_block(L - 1))
        sums, sums64 = _sum_across_ranks(sums, sums64, group, cfg.get("peers"), L - 1)
        du, pooled, dx = None, demb, None
        share = 1.0 / _world(group)
        for l in range(L - 1, -1, -1):
            t_in, act_in, z, W, scale, mean, rstd, agg = saved[l]
            need_du = l > 0 or ctx.x_needs_grad
            prev_mean = saved[l - 1][5] if l > 0 else None
            prev_rstd = saved[l - 1][6] if l > 0 else None
            bn = BnBwd(scale, mean, rstd, sums, count, training, sums64)
            out = (flat.view(layer_params[l][0]), flat.view(layer_params[l][1])) if flat is not None else None
            dW, db, du_in, prev_sums, prev_sums64 = eng.layer_bwd(kind, du, pooled, z, act_out, bn, t_in, act_in, W, csr, ptr, B,
                                                     need_du, prev_mean, prev_rstd, agg, out=out,
                                                     prev_out=bn_block(l - 1) if l > 0 else None)
            # d gamma = sum dy*xhat, d beta = sum dy.  Under data parallelism `sums` is already the global sum while
            # every other gradient is this rank's share and Trainer adds the ranks up: hand out 1/world of it.
            if flat is not None:
                if share != 1.0:
                    sums.mul_(share)          # in place, after the kernels that read the sums have been enqueued
                grads[4 * l + 0], grads[4 * l + 1] = flat.view(layer_params[l][0]), flat.view(layer_params[l][1])
                grads[4 * l + 2], grads[4 * l + 3] = flat.view(layer_params[l][2]), flat.view(layer_params[l][3])
            else:
                grads[4 * l + 0], grads[4 * l + 1] = dW, db
                grads[4 * l + 2] = sums[1] if share == 1.0 else sums[1] * share
                grads[4 * l + 3] = sums[0] if share == 1.0 else sums[0] * share
            sums, sums64 = _sum_across_ranks(prev_sums, prev_sums64, group, cfg.get("peers"), l - 1)
            du, pooled, act_out = du_in, None, act_in
            if l == 0:
                dx = du_in
        return (None, dx if ctx.x_needs_grad else None, *grads)


class _HeadFn(torch.autograd.Function):
    """``logits = W1 dropout(relu(W0 emb + b0)) + b1`` (reference ``models.py:196-201``)."""

    @staticmethod
    def forward(ctx, cfg: dict, emb, W0, b0, W1, b1):
        eng = cfg["engine"]
        emb, W0, b0, W1, b1 = (q.contiguous() for q in (emb, W0, b0, W1, b1))
        p_eff = cfg["dropout"] if cfg["training"] else 0.0
        hidden, logits = eng.head_fwd(emb, W0, b0, W1, b1, p_eff, cfg["seed"], cfg["graph_base"],
                                      cfg.get("salt") if p_eff > 0.0 else None)
        ctx.cfg, ctx.p_eff = cfg, p_eff
        ctx.save_for_backward(emb, hidden, W0, W1)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        emb, hidden, W0, W1 = ctx.saved_tensors
        flat, hp = ctx.cfg.get("flat"), ctx.cfg.get("head_params")
        out = tuple(flat.view(q) for q in hp) if flat is not None else None
        demb, dW0, db0, dW1, db1 = ctx.cfg["engine"].head_bwd(emb, hidden, dlogits.contiguous(), W0, W1, ctx.p_eff, out=out)
        if flat is not None:
            dW0, db0, dW1, db1 = (flat.view(q) for q in hp)       # fresh views: autograd adopts them as .grad without a copy
        return None, demb, dW0, db0, dW1, db1


# ---------------------------------------------------------------------------
# parameter holders
# ---------------------------------------------------------------------------

class GCNLayer(nn.Module):
    """Parameters of one GCN layer: ``linear.weight [out, in]`` (no bias), ``bias [out]``
    (reference ``models.py:78-82``).  The convolution itself runs inside the model's fused path."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.linear = nn.Linear(in_channels, out_channels, bias=False)
        self.bias = nn.Parameter(torch.zeros(out_channels))
        nn.init.xavier_uniform_(self.linear.weight)

    def tensors(self):
        return self.linear.weight, self.bias

    def forward(self, x, edge_index, edge_weight):
        """Stand-alone ``A^ (x W^T) + b`` over one block of nodes (all nodes form one subject)."""
        return _single_layer("gcn", self, x, edge_index, edge_weight)


class SAGELayer(nn.Module):
    """Parameters of one GraphSAGE layer: ``linear.weight [out, 2*in]``, ``linear.bias [out]``
    (reference ``models.py:130-134``)."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.linear = nn.Linear(in_channels * 2, out_channels)
        nn.init.xavier_uniform_(self.linear.weight)

    def tensors(self):
        return self.linear.weight, self.linear.bias

    def forward(self, x, edge_index, edge_weight):
        """Stand-alone ``relu([x || mean_w(x_nbrs)] W^T + b)`` over one block of nodes."""
        return _single_layer("sage", self, x, edge_index, edge_weight)


def _single_layer(kind: str, layer: nn.Module, x, edge_index, edge_weight):
    """Inference-only convenience for calling a layer by itself (not differentiable)."""
    if torch.is_grad_enabled() and any(q.requires_grad for q in (x, *layer.parameters())):
        raise RuntimeError("stand-alone layer calls are inference-only; wrap in torch.no_grad() "
                           "or differentiate through GCNConnectome / GraphSAGEConnectome")
    dev = x.device if x.is_cuda else _engine.default_device()
    x, edge_index, edge_weight = x.to(dev), edge_index.to(dev), edge_weight.to(dev)
    n = x.shape[0]
    ptr = torch.tensor([0, n], dtype=torch.int64, device=dev)
    batch = ConnectomeBatch(x, edge_index, edge_weight, torch.zeros(n, dtype=torch.int64, device=dev), None, ptr)
    csr = batch.ensure_csr()
    eng = _engine.engine_for(x)
    W, b = (q.detach().to(dev).contiguous() for q in layer.tensors())
    z, _, _ = eng.layer_fwd(kind, x.contiguous().float(), Act(), W, b, csr, ptr, 1, want_stats=False)
    return z


class _ConnectomeClassifier(nn.Module):
    """Shared machinery of the two classifiers (reference ``models.py:159-216`` / ``219-266``)."""

    kind = ""
    layer_cls = None

    def __init__(self, in_channels: int, hidden_dim: int = 64, num_classes: int = 2, num_layers: int = 3,
                 dropout: float = 0.3):
        super().__init__()
        self.dropout = dropout
        widths = [in_channels] + [hidden_dim] * num_layers
        self.convs = nn.ModuleList([self.layer_cls(a, b) for a, b in zip(widths[:-1], widths[1:])])
        self.batch_norms = nn.ModuleList([nn.BatchNorm1d(hidden_dim) for _ in range(num_layers)])
        self.classifier = nn.Sequential(
            nn.Linear(hidden_dim, hidden_dim // 2),
            nn.ReLU(),
            nn.Dropout(dropout),
            nn.Linear(hidden_dim // 2, num_classes),
        )
        self.process_group = None   # torch.distributed group for SyncBN statistics (None = default, False = no collectives)
        # Inference path: "auto" runs cgnn_eval_fused_fwd (whole network in one kernel, two launches) when the batch is small
        # enough to be launch-latency bound - at most FUSED_EVAL_MAX_ROWS nodes - and layer by layer otherwise: a unit's
        # layers are serial inside the fused kernel, while the per-layer kernels overlap the phases of different units
        # (measured on B200, 4096 x 360-node subjects: 1.29 ms fused, 0.91 ms layer by layer).  True / False force one.
        self.fused_eval = "auto"
        self._salt = None           # device dropout salt of a CUDA-graphed step (Trainer.capture), else None
        self._graph_seed = 0
        self._flat = None           # flat parameter / gradient buffers (Trainer.enable_fused_step), else None

    # -- plumbing --------------------------------------------------------------------------
    def _ready(self, batch: ConnectomeBatch) -> ConnectomeBatch:
        """Batch on a CUDA device with its CSR; parameters follow the batch onto that device."""
        if not batch.node_features.is_cuda:
            batch = batch.to(_engine.default_device())
        batch.ensure_csr()
        dev = batch.node_features.device
        if next(self.parameters()).device != dev:
            self.to(dev)   # in place: Parameter identity (and any optimizer built earlier) is preserved
        return batch

    def _cfg(self, batch: ConnectomeBatch) -> dict:
        training = self.training
        need_seed = training and self.dropout > 0.0
        cfg = dict(engine=_engine.engine_for(batch.node_features), batch=batch, kind=self.kind, training=training,
                   dropout=float(self.dropout), bns=list(self.batch_norms), group=self.process_group,
                   graph_base=batch.graph_base, salt=self._salt, flat=self._flat)
        # data parallel on an NVLink box: BatchNorm statistics and their backward sums travel through peer memory in one
        # kernel per exchange (peers.py) instead of NCCL collectives; `model.peer_collectives = False` keeps NCCL
        cfg["peers"] = None
        if training and getattr(self, "peer_collectives", True) and _world(self.process_group) > 1:
            from .peers import peer_exchange_for
            cfg["peers"] = peer_exchange_for(self.process_group, batch.node_features.device)
        # a graphed step keeps its dropout stream on the device (salt words refreshed inside the graph): the host seed is fixed
        cfg["seed"] = (self._graph_seed if self._salt is not None else _draw_seed(batch.node_features.device)) if need_seed else 0
        if self._flat is not None:
            cfg["layer_params"] = [(*conv.tensors(), bn.weight, bn.bias) for conv, bn in zip(self.convs, self.batch_norms)]
            fc0, fc1 = self.classifier[0], self.classifier[3]
            cfg["head_params"] = (fc0.weight, fc0.bias, fc1.weight, fc1.bias)
        return cfg

    def _encode(self, batch: ConnectomeBatch, cfg: dict) -> torch.Tensor:
        params = []
        for conv, bn in zip(self.convs, self.batch_norms):
            params += [*conv.tensors(), bn.weight, bn.bias]
        # decided here (inside Function.forward grad mode is always off): under torch.no_grad() / with everything frozen no
        # backward can follow, and the forward may skip what only a backward pass would read
        cfg["forward_only"] = not (torch.is_grad_enabled() and (batch.node_features.requires_grad or
                                                                any(q is not None and q.requires_grad for q in params)))
        return _EncodeFn.apply(cfg, batch.node_features, *params)

    def _fused_eval(self, batch: ConnectomeBatch, want_logits: bool):
        """Inference fast path (``cgnn_eval_fused_fwd``): eval mode, no autograd graph wanted - the whole network, readout
        and head in ONE kernel, activations never leave the SM.  None when not applicable / not covered."""
        if not self.fused_eval or (self.fused_eval == "auto" and batch.num_nodes > FUSED_EVAL_MAX_ROWS):
            return None
        if self.training or (torch.is_grad_enabled() and (batch.node_features.requires_grad or
                                                          any(p.requires_grad for p in self.parameters()))):
            return None
        layers = []
        for conv, bn in zip(self.convs, self.batch_norms):
            W, b = conv.tensors()
            if bn.running_mean is None or bn.running_var is None:
                return None
            layers.append((W.detach().contiguous(), b.detach().contiguous(),
                           None if bn.weight is None else bn.weight.detach().contiguous(),
                           None if bn.bias is None else bn.bias.detach().contiguous(),
                           bn.running_mean.contiguous(), bn.running_var.contiguous(), bn.eps))
        fc0, fc1 = self.classifier[0], self.classifier[3]
        head = tuple(t.detach().contiguous() for t in (fc0.weight, fc0.bias, fc1.weight, fc1.bias))
        eng = _engine.engine_for(batch.node_features)
        return eng.eval_fused(self.kind, batch.node_features.contiguous(), layers, head, batch.csr, batch.ptr,
                              batch.num_graphs, want_logits)

    # -- public API ------------------------------------------------------------------------
    def encode(self, batch: ConnectomeBatch) -> torch.Tensor:
        """Graph-level embeddings ``[B, hidden_dim]``."""
        batch = self._ready(batch)
        fused = self._fused_eval(batch, want_logits=False)
        if fused is not None:
            return fused[0]
        return self._encode(batch, self._cfg(batch))

    def forward(self, batch: ConnectomeBatch) -> torch.Tensor:
        """Class logits ``[B, num_classes]``."""
        batch = self._ready(batch)
        fused = self._fused_eval(batch, want_logits=True)
        if fused is not None:
            return fused[1]
        cfg = self._cfg(batch)
        emb = self._encode(batch, cfg)
        fc0, fc1 = self.classifier[0], self.classifier[3]
        return _HeadFn.apply(cfg, emb, fc0.weight, fc0.bias, fc1.weight, fc1.bias)


class GCNConnectome(_ConnectomeClassifier):
    """L x (GCN conv -> BatchNorm -> ReLU -> dropout) -> mean-pool -> MLP (reference ``models.py:159-216``)."""
    kind = "gcn"
    layer_cls = GCNLayer


class GraphSAGEConnectome(_ConnectomeClassifier):
    """L x (SAGE conv incl. ReLU -> BatchNorm -> dropout) -> mean-pool -> MLP (reference ``models.py:219-266``)."""
    kind = "sage"
    layer_cls = SAGELayer
