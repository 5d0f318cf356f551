"""Subject graphs, device-resident subject stores and batch collation.

Host-side mirror of reference ``connectome_gnn/graph.py`` with the same public names and
signatures (``ConnectomeGraph``, ``ConnectomeBatch``, ``collate_graphs``, ``ConnectomeDataLoader``).
What changes is where the work happens: subjects are packed once into a device arena
(:class:`SubjectStore`); a batch is then produced by ONE kernel sequence
(``cgnn_collate_csr``) that gathers the selected subjects, writes the six reference fields
bit-exactly as reference ``graph.py:143-167`` would, and builds the per-subject CSR, weighted
degrees and GCN normalisation that reference ``models.py:94-108,147-148`` recompute in every
layer.  Nothing here has a CPU compute path; host code only packs, indexes and shards.
"""

from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Sequence

import contextlib

import numpy as np
import torch

from . import _engine
from ._lib import StoreT

__all__ = ["ConnectomeGraph", "ConnectomeBatch", "BatchCSR", "SubjectStore", "collate_graphs",
           "ConnectomeDataLoader", "shard_bounds"]


# ---------------------------------------------------------------------------
# One subject
# ---------------------------------------------------------------------------

@dataclass
class ConnectomeGraph:
    """One subject: regions are nodes, weighted connections are directed COO edges (both
    directions present for an undirected connectome).  Same fields as reference ``graph.py:27-56``."""

    node_features: torch.Tensor            # [N, F] float32
    edge_index: torch.Tensor               # [2, E] int64, subject-local node ids
    edge_weight: torch.Tensor              # [E] float32
    label: Optional[torch.Tensor] = None   # scalar
    subject_id: str = "unknown"

    @property
    def num_nodes(self) -> int:
        return int(self.node_features.shape[0])

    @property
    def num_edges(self) -> int:
        return int(self.edge_index.shape[1])

    @property
    def num_features(self) -> int:
        return int(self.node_features.shape[1])

    def adjacency_matrix(self) -> torch.Tensor:
        """Dense ``[N, N]`` weight matrix (later duplicates overwrite earlier ones)."""
        n = self.num_nodes
        dense = self.edge_weight.new_zeros((n, n))
        dense[self.edge_index[0], self.edge_index[1]] = self.edge_weight
        return dense

    def degree(self) -> torch.Tensor:
        """Weighted out-degree ``[N]``."""
        out = self.edge_weight.new_zeros(self.num_nodes)
        return out.scatter_add_(0, self.edge_index[0], self.edge_weight)

    def to(self, device) -> "ConnectomeGraph":
        move = lambda t: None if t is None else t.to(device)
        return ConnectomeGraph(move(self.node_features), move(self.edge_index), move(self.edge_weight),
                               move(self.label), self.subject_id)


# ---------------------------------------------------------------------------
# Batch
# ---------------------------------------------------------------------------

class _FullCollate:
    """What a lean batch and its lean CSR share to fill themselves in: the store and the subject indices, and - once
    somebody asked - the full batch collated from them.  Holds no reference to the lean objects (no cycle)."""
    __slots__ = ("store", "ids_np", "ids_device", "full")

    def __init__(self, store, ids_np, ids_device):
        self.store, self.ids_np, self.ids_device, self.full = store, ids_np, ids_device, None

    def get(self):
        if self.full is None:
            self.full = self.store.collate(self.ids_np, ids_device=self.ids_device)
        return self.full


class BatchCSR:
    """Device CSR of a batch (``cgnn_csr_t``): by-destination and by-source rows in stable COO
    order, raw and GCN-normalised weights, D^ / d^-1/2 / w_sum.  See include/cgnn.h.

    A LEAN instance (what ``SubjectStore.collate(prepare_for=...)`` returns) carries only ``graph_meta``, ``eptr`` and
    the packed aggregation blobs in ``agg`` - all the tensor-core layer kernels read.  The eleven arrays are
    materialised on first attribute access by collating the same subjects again with every output requested
    (``materialize``); ``peek`` looks without triggering that."""

    _ARRAYS = ("in_rowptr", "in_col", "in_w", "in_wn", "out_rowptr", "out_col", "out_w", "out_wn", "deg", "dinv", "wsum")
    _FIELDS = _ARRAYS + ("graph_meta", "eptr")

    def __init__(self, in_rowptr=None, in_col=None, in_w=None, in_wn=None, out_rowptr=None, out_col=None, out_w=None,
                 out_wn=None, deg=None, dinv=None, wsum=None, graph_meta=None, eptr=None, max_nodes: int = 0,
                 max_edges: int = 0, agg: Optional[dict] = None, num_edges: Optional[int] = None, fill=None):
        d = self.__dict__
        for k, v in zip(self._ARRAYS, (in_rowptr, in_col, in_w, in_wn, out_rowptr, out_col, out_w, out_wn, deg, dinv, wsum)):
            d["_" + k] = v
        self.graph_meta = graph_meta     # [B, 4] int32 {first row, rows, first edge, edges} per subject
        self.eptr = eptr                 # [B+1] int64 edge prefix sums
        self.max_nodes = max_nodes       # largest subject of the batch (sizes shared-memory tiles)
        self.max_edges = max_edges       # most edges in one subject
        self.agg = agg                   # model family -> (agg_in, agg_out | None, row_graph) packed blobs
        self.num_edges = int(num_edges) if num_edges is not None else (int(in_col.shape[0]) if in_col is not None else 0)
        self._fill = fill                # lean instances: _FullCollate of the same subjects (holds no reference back)

    def peek(self, name: str):
        """The array if it exists on the device, else None (never materialises)."""
        if name in self._ARRAYS:
            return self.__dict__["_" + name]
        return getattr(self, name)

    def is_full(self) -> bool:
        return self.__dict__["_in_col"] is not None

    def materialize(self) -> "BatchCSR":
        """Make sure the CSR arrays exist (a lean batch collates its subjects again with every output requested)."""
        if not self.is_full():
            if self._fill is None:
                raise RuntimeError("this BatchCSR has no CSR arrays and no way to build them")
            self._adopt(self._fill.get().csr)
        return self

    def _adopt(self, full: "BatchCSR") -> None:
        for k in self._ARRAYS:
            self.__dict__["_" + k] = full.peek(k)
        self._fill = None

    def __getattr__(self, name: str):
        # only reached for names that are not instance attributes: the lazily materialised arrays
        if name in BatchCSR._ARRAYS:
            self.materialize()
            return self.__dict__["_" + name]
        raise AttributeError(name)

    def to(self, device) -> "BatchCSR":
        self.materialize()
        mv = lambda t: None if t is None else t.to(device)
        return BatchCSR(*[mv(self.peek(f)) for f in self._FIELDS], max_nodes=self.max_nodes, max_edges=self.max_edges,
                        num_edges=self.num_edges)


@dataclass
class ConnectomeBatch:
    """Block-diagonal packing of B subjects; the first six fields are exactly the reference's
    (``graph.py:101-122``).  ``csr`` carries the device CSR the kernels read; ``row_base`` /
    ``graph_base`` / ``global_num_graphs`` describe this rank's slice of a data-parallel batch."""

    node_features: torch.Tensor            # [sum N, F]
    edge_index: torch.Tensor               # [2, sum E] int64, node ids offset per subject
    edge_weight: torch.Tensor              # [sum E]
    batch: torch.Tensor                    # [sum N] int64 subject index of each node
    labels: Optional[torch.Tensor]         # [B] int64
    ptr: torch.Tensor                      # [B+1] int64
    csr: Optional[BatchCSR] = None
    row_base: int = 0
    graph_base: int = 0
    global_num_graphs: Optional[int] = None
    global_num_nodes: Optional[int] = None
    # lean batches (SubjectStore.collate(prepare_for=...)): edge_index / edge_weight / batch are None until somebody
    # reads them; `_fill` (a _FullCollate, shared with the CSR) then collates the same subjects again with every
    # reference field requested.  It must not refer back to the batch: a reference cycle would keep every batch (GBs of
    # device memory) alive until the cyclic garbage collector happens to run
    _fill: Optional[object] = None

    _LAZY = ("edge_index", "edge_weight", "batch")

    def __getattribute__(self, name):
        v = object.__getattribute__(self, name)
        if v is None and name in ConnectomeBatch._LAZY:
            fill = object.__getattribute__(self, "_fill")
            if fill is not None:
                full = fill.get()
                for k in ConnectomeBatch._LAZY:
                    object.__setattr__(self, k, object.__getattribute__(full, k))
                object.__setattr__(self, "_fill", None)
                v = object.__getattribute__(self, name)
        return v

    @property
    def num_graphs(self) -> int:
        return int(self.ptr.shape[0]) - 1

    @property
    def num_nodes(self) -> int:
        return int(self.node_features.shape[0])

    @property
    def device(self) -> torch.device:
        return self.node_features.device

    def to(self, device) -> "ConnectomeBatch":
        device = torch.device(device)
        if device.type == "cuda" and device.index is None and self.node_features.is_cuda:
            return self
        if device == self.node_features.device:
            return self
        move = lambda t: None if t is None else t.to(device)
        return ConnectomeBatch(   # reading the lazy fields materialises them: a moved batch is a full batch
            move(self.node_features), move(self.edge_index), move(self.edge_weight), move(self.batch),
            move(self.labels), move(self.ptr), None if self.csr is None else self.csr.to(device),
            self.row_base, self.graph_base, self.global_num_graphs, self.global_num_nodes)

    def ensure_csr(self) -> BatchCSR:
        """Build the CSR for a batch that was assembled by hand or moved from the host
        (COO grouped by subject, as ``collate_graphs`` emits it)."""
        if self.csr is None:
            eng = _engine.engine_for(self.node_features)
            counts = (self.ptr[1:] - self.ptr[:-1])
            max_nodes = int(counts.max().item()) if counts.numel() else 0
            ei = self.edge_index.contiguous()
            if ei.shape[1] > 0:
                # the CSR builder walks the COO list subject by subject: edges must be grouped by subject (as
                # collate_graphs emits them) and stay inside their subject - anything else would silently drop edges
                gs, gd = self.batch[ei[0]], self.batch[ei[1]]
                if not bool((gs == gd).all()) or not bool((gs[1:] >= gs[:-1]).all()):
                    raise ValueError("ConnectomeBatch.edge_index must list edges grouped by subject, both endpoints in "
                                     "the same subject (sort the COO list by batch[edge_index[0]] first)")
            csr, eptr = eng.csr_from_coo(ei, self.edge_weight.contiguous(), self.ptr.contiguous(), self.num_graphs,
                                         self.num_nodes, int(ei.shape[1]), max_nodes)
            max_edges = int((eptr[1:] - eptr[:-1]).max().item()) if self.num_graphs else 0
            self.csr = BatchCSR(**csr, eptr=eptr, max_nodes=max_nodes, max_edges=max_edges)
        return self.csr


# ---------------------------------------------------------------------------
# Subject store: the dataset packed once
# ---------------------------------------------------------------------------

def _is_pair_list(ei: torch.Tensor, w: torch.Tensor, n_edges: np.ndarray) -> bool:
    """True when every subject's COO list is a sequence of adjacent (s -> d), (d -> s) pairs of equal weight."""
    if ei.shape[1] == 0 or bool((n_edges % 2).any()):
        return False
    a, b = ei[:, 0::2], ei[:, 1::2]
    return bool(torch.equal(a[0], b[1]) and torch.equal(a[1], b[0]) and torch.equal(w[0::2], w[1::2]))


def pack_graphs(graphs: Sequence[ConnectomeGraph], compact: Optional[bool] = None, pairs: Optional[bool] = None) -> dict:
    """Host packing of a list of subjects into arena arrays (pure indexing, no arithmetic).

    Returns CPU tensors: ``x [sum N, F]`` f32, ``src/dst [sum E]`` int32 subject-local ids,
    ``w [sum E]`` f32, ``node_ptr/edge_ptr [S+1]`` int64, ``label [S]`` int64 (0 where absent)
    and ``has_label [S]`` bool.  ``compact`` packs both endpoints into one int32; ``pairs`` (with ``compact``)
    additionally stores one entry per undirected edge when every subject's list allows it (``edge_pairs`` = 1 in the
    result, else 0 and the plain compact form).  Both default to "when possible" (``None``): subjects under 65536 nodes,
    lists made of adjacent reversed pairs - the smallest arena, and the collate kernel then sorts once for both
    directions; pass ``False`` for the plain forms.  Raises ``ValueError`` on ragged feature widths or edge
    endpoints outside ``[0, N_s)`` - the kernels index shared-memory tiles with them."""
    if len(graphs) == 0:
        raise ValueError("cannot pack an empty list of ConnectomeGraph")
    feat = graphs[0].num_features
    n_nodes = np.fromiter((g.num_nodes for g in graphs), dtype=np.int64, count=len(graphs))
    n_edges = np.fromiter((g.num_edges for g in graphs), dtype=np.int64, count=len(graphs))
    if any(g.num_features != feat for g in graphs):
        raise ValueError("all subjects must have the same number of node features")
    node_ptr = np.zeros(len(graphs) + 1, dtype=np.int64)
    edge_ptr = np.zeros(len(graphs) + 1, dtype=np.int64)
    np.cumsum(n_nodes, out=node_ptr[1:])
    np.cumsum(n_edges, out=edge_ptr[1:])

    cpu = lambda t: t.detach().to("cpu")
    x = torch.cat([cpu(g.node_features).to(torch.float32) for g in graphs], dim=0).contiguous()
    ei = torch.cat([cpu(g.edge_index).to(torch.int64) for g in graphs], dim=1)
    w = torch.cat([cpu(g.edge_weight).to(torch.float32) for g in graphs], dim=0).contiguous()
    if ei.numel():
        bound = torch.from_numpy(np.repeat(n_nodes, n_edges))
        if int(ei.min()) < 0 or bool((ei >= bound.unsqueeze(0)).any()):
            raise ValueError("edge_index refers to a node outside its subject")
    has_label = np.fromiter((g.label is not None for g in graphs), dtype=bool, count=len(graphs))
    label = torch.zeros(len(graphs), dtype=torch.int64)
    if has_label.any():
        vals = torch.stack([cpu(g.label).reshape(()).to(torch.int64) for g in graphs if g.label is not None])
        label[torch.from_numpy(np.nonzero(has_label)[0])] = vals
    edge_pairs = 0
    if compact is None:
        compact = not (n_nodes.size and int(n_nodes.max()) > 65535)
    if pairs is None:
        pairs = bool(compact)
    if compact:
        # both endpoints of an edge in one int32 (src | dst << 16): 4 bytes less per edge over PCIe
        if n_nodes.size and int(n_nodes.max()) > 65535:
            raise ValueError("compact packing needs every subject to have < 65536 nodes")
        src = (ei[0] | (ei[1] << 16)).to(torch.int64)
        src = torch.where(src >= 2 ** 31, src - 2 ** 32, src).to(torch.int32).contiguous()
        dst = torch.zeros(0, dtype=torch.int32)
        # undirected connectomes list every edge as two adjacent directed edges of equal weight (the reference
        # generator and its real-data recipe both do): one entry per pair then carries the same information
        if pairs and _is_pair_list(ei, w, n_edges):
            src, w, edge_pairs = src[0::2].contiguous(), w[0::2].contiguous(), 1
    else:
        src, dst = ei[0].to(torch.int32).contiguous(), ei[1].to(torch.int32).contiguous()
    return dict(
        x=x, src=src, dst=dst, w=w,
        node_ptr=torch.from_numpy(node_ptr), edge_ptr=torch.from_numpy(edge_ptr), label=label,
        has_label=has_label, num_features=feat, edge_pairs=edge_pairs)


class SubjectStore:
    """All subjects of a dataset, packed once and resident on one GPU (``cgnn_store_t``).

    ``collate(ids)`` turns any selection of subject indices into a :class:`ConnectomeBatch`
    entirely on the device; the only per-batch host->device traffic is ``ids`` (8 B / subject).
    """

    def __init__(self, packed: dict, device=None):
        # default_device() raises when no CUDA device is visible; collate() goes through
        # engine_for(), which refuses non-CUDA tensors - there is no CPU path.
        self.device = torch.device(device) if device is not None else _engine.default_device()
        self.num_features = int(packed["num_features"])
        self.node_ptr_host = packed["node_ptr"].numpy()
        self.edge_ptr_host = packed["edge_ptr"].numpy()
        self.has_label = np.asarray(packed["has_label"], dtype=bool)
        nb = lambda t: t.to(self.device, non_blocking=True)
        self.x, self.src, self.dst, self.w = nb(packed["x"]), nb(packed["src"]), nb(packed["dst"]), nb(packed["w"])
        self.node_ptr, self.edge_ptr = nb(packed["node_ptr"]), nb(packed["edge_ptr"])
        self.label = nb(packed["label"])
        compact = self.dst.numel() == 0 and self.src.numel() > 0     # pack_graphs(compact=True): src | dst << 16
        self.edge_pairs = int(packed.get("edge_pairs", 0))
        self._struct = StoreT(self.x.data_ptr(), self.src.data_ptr(), None if compact else self.dst.data_ptr(), self.w.data_ptr(),
                              self.node_ptr.data_ptr(), self.edge_ptr.data_ptr(), self.label.data_ptr(),
                              self.num_features, self.edge_pairs)

    def _upload_ids(self, ids_np: np.ndarray) -> torch.Tensor:
        """Subject indices to the device through a small ring of PINNED staging buffers, fetched by a kernel."""
        n = int(ids_np.size)
        if self.device.type != "cuda":       # test-only simulator engine: host pointers
            return torch.from_numpy(ids_np).to(self.device)
        ring = getattr(self, "_ids_ring", None)
        if ring is None or ring[0][0].numel() < n:
            ring = self._ids_ring = [[torch.empty(max(n, 1), dtype=torch.int64).pin_memory(), None] for _ in range(4)]
            self._ids_next = 0
        slot = ring[self._ids_next % len(ring)]
        self._ids_next += 1
        if slot[1] is not None:
            slot[1].synchronize()            # the copy that last used this buffer (four collates ago) has long finished
        slot[0][:n].copy_(torch.from_numpy(ids_np))
        # a kernel reads the pinned buffer over the bus (cgnn_fetch_ids): a DMA of these few KB would wait on the copy
        # engine behind a StreamingStore upload of tens of MB
        eng = _engine.engine_for(self.x)
        out = torch.empty(n, dtype=torch.int64, device=self.device)
        eng._call("cgnn_fetch_ids", slot[0].data_ptr(), n, out.data_ptr(), eng.stream())
        slot[1] = torch.cuda.Event()
        slot[1].record(torch.cuda.current_stream(self.device))
        return out

    def release(self) -> None:
        """Mark the work enqueued so far on the current stream as the last use of this arena's contents: a later
        ``reload`` on another stream waits for it on the device, so the host never has to synchronise."""
        self._free = torch.cuda.Event()
        self._free.record(torch.cuda.current_stream(self.device))

    @classmethod
    def from_graphs(cls, graphs: Sequence[ConnectomeGraph], device=None) -> "SubjectStore":
        return cls(pack_graphs(graphs), device)

    def reload(self, packed: dict, stream: Optional["torch.cuda.Stream"] = None) -> "SubjectStore":
        """Overwrite this store's device arenas with another packed set of the same shapes (host tensors, ideally
        pinned), enqueued on ``stream`` (default: the current stream).  ``collate`` makes the consuming stream wait
        for the upload, so a side stream overlaps the copy with whatever the compute stream is still doing - the
        building block of :class:`StreamingStore`."""
        names = ("x", "src", "dst", "w", "node_ptr", "edge_ptr", "label")
        for k in names:
            if tuple(packed[k].shape) != tuple(getattr(self, k).shape) or packed[k].dtype != getattr(self, k).dtype:
                raise ValueError(f"reload: '{k}' differs in shape or dtype from the resident arena")
        if int(packed.get("edge_pairs", 0)) != self.edge_pairs:
            raise ValueError("reload: the packed set and the resident arena differ in their edge layout (edge_pairs)")
        self.node_ptr_host = packed["node_ptr"].numpy()
        self.edge_ptr_host = packed["edge_ptr"].numpy()
        self.has_label = np.asarray(packed["has_label"], dtype=bool)
        ctx = torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()
        with ctx:
            free = getattr(self, "_free", None)
            if free is not None:       # the previous consumer of this arena (release()) must have finished with it
                torch.cuda.current_stream(self.device).wait_event(free)
            for k in names:
                getattr(self, k).copy_(packed[k], non_blocking=True)
            self._ready = torch.cuda.Event()
            self._ready.record()
        return self

    def __len__(self) -> int:
        return int(self.node_ptr_host.shape[0]) - 1

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.x, self.src, self.dst, self.w, self.node_ptr,
                                                          self.edge_ptr, self.label))

    def collate(self, ids, *, row_base: int = 0, graph_base: int = 0, global_num_graphs: Optional[int] = None,
                global_num_nodes: Optional[int] = None, ids_device: Optional[torch.Tensor] = None,
                prepare_for: Optional[str] = None, lean: Optional[bool] = None, backward: bool = True) -> ConnectomeBatch:
        """Device collate of subjects ``ids`` (host int sequence / array, in batch order).  ``prepare_for`` ("gcn" /
        "sage") makes the collate kernel emit that model family's packed aggregation structure in the same pass
        instead of a separate launch at the first layer (same bits either way) - and, unless ``lean=False``, makes
        the batch LEAN: the kernel writes only what that model's kernels read (node features, labels, ``ptr``, the
        packed structure); ``edge_index`` / ``edge_weight`` / ``batch`` and the CSR arrays are filled in - bit for bit
        what a full collate writes - the first time they are read.  ``backward=False`` (inference) also skips the
        by-source half of the structure."""
        ids_np = np.asarray(ids, dtype=np.int64).reshape(-1)
        n_sel = self.node_ptr_host[ids_np + 1] - self.node_ptr_host[ids_np]
        e_sel = self.edge_ptr_host[ids_np + 1] - self.edge_ptr_host[ids_np]
        rows, edges = int(n_sel.sum()), int(e_sel.sum())
        max_nodes = int(n_sel.max()) if ids_np.size else 0
        max_edges = int(e_sel.max()) if ids_np.size else 0
        labelled = self.has_label[ids_np]
        all_labelled = bool(labelled.all()) and ids_np.size > 0
        if ids_device is None:
            ids_device = self._upload_ids(ids_np)
        ready = getattr(self, "_ready", None)
        if ready is not None:      # an upload enqueued on another stream (reload): order this stream after it
            torch.cuda.current_stream(self.device).wait_event(ready)
        eng = _engine.engine_for(self.x)
        want_lean = (prepare_for is not None) if lean is None else (bool(lean) and prepare_for is not None)
        out, csr, blobs = eng.collate_csr(self._struct, ids_device, int(ids_np.size), rows, edges, max_nodes, max_edges,
                                          self.num_features, all_labelled, prepare_for, lean=want_lean,
                                          need_out=backward)
        labels = out["labels"]
        if ids_np.size == 0:         # an empty slice of a data-parallel batch: no subjects, no labels (not "unlabelled")
            labels = torch.empty(0, dtype=torch.int64, device=self.device)
        if not all_labelled and labelled.any():
            # reference quirk (graph.py:155-156,165): only labelled subjects contribute, so the
            # stack is shorter than B
            labels = self.label[torch.from_numpy(ids_np[labelled]).to(self.device)]
        bcsr = BatchCSR(**csr, eptr=out["eptr"], max_nodes=max_nodes, max_edges=max_edges, num_edges=edges)
        if blobs is not None:
            bcsr.agg = {prepare_for: blobs}
        batch = ConnectomeBatch(
            out["node_features"], out["edge_index"], out["edge_weight"], out["batch"], labels, out["ptr"],
            bcsr, row_base, graph_base, global_num_graphs, global_num_nodes)
        if out["edge_index"] is None:          # lean: the missing fields are one full collate of the same subjects away
            src = _FullCollate(self, ids_np, ids_device)
            object.__setattr__(batch, "_fill", src)
            bcsr._fill = src
        return batch


class StreamingStore:
    """Subjects that live in (pinned) host memory and visit the GPU one packed set at a time: ``depth`` device arenas, the
    upload of set k+1 runs on a side stream while set k is being collated and processed (pinned-host
    staging for datasets that do not stay resident).  Usage::

        ss = StreamingStore(packed_sets[0], device)
        ss.prefetch(packed_sets[0])
        for k in range(len(packed_sets)):
            store = ss.next()                       # set k, upload ordered before this stream's next kernels
            if k + 1 < len(packed_sets):
                ss.prefetch(packed_sets[k + 1])     # starts now, overlaps the work below
            batch = store.collate(ids)
            ...

    All sets must have the shapes of the first one (same subjects-per-set layout).  Call ``store.release()`` after the
    last kernel that reads a set has been enqueued (or synchronise with the host): ``prefetch`` into that arena then
    waits for it on the device."""

    def __init__(self, packed: dict, device=None, depth: int = 2):
        self.device = torch.device(device) if device is not None else _engine.default_device()
        self.stream = torch.cuda.Stream(device=self.device)
        # `depth` device arenas: uploads run up to depth - 1 sets ahead of the set being processed (depth 3 gives every
        # upload two processing steps of slack - short inference steps no longer wait for the host link)
        self._arenas = [SubjectStore(packed, self.device) for _ in range(max(2, int(depth)))]
        torch.cuda.current_stream(self.device).synchronize()
        self._filled, self._taken = 0, 0

    def prefetch(self, packed: dict) -> None:
        if self._filled - self._taken >= len(self._arenas):
            raise RuntimeError("StreamingStore.prefetch(): every arena holds a set that has not been taken yet")
        self._arenas[self._filled % len(self._arenas)].reload(packed, self.stream)
        self._filled += 1

    def next(self) -> SubjectStore:
        if self._taken >= self._filled:
            raise RuntimeError("StreamingStore.next() without a matching prefetch()")
        store = self._arenas[self._taken % len(self._arenas)]
        self._taken += 1
        return store


def collate_graphs(graphs: list) -> ConnectomeBatch:
    """Collate a list of :class:`ConnectomeGraph` into a device-resident :class:`ConnectomeBatch`
    (same signature as reference ``graph.py:143``).  The list is packed on the host, shipped in one
    copy per array and collated by the device kernel; results live on the current CUDA device."""
    store = SubjectStore.from_graphs(graphs)
    return store.collate(np.arange(len(graphs), dtype=np.int64))


# ---------------------------------------------------------------------------
# Loader
# ---------------------------------------------------------------------------

def shard_bounds(count: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous slice ``[lo, hi)`` of a chunk of ``count`` subjects owned by ``rank``.
    Slices differ by at most one subject; ranks past the end of a short chunk get an empty slice."""
    base, extra = divmod(count, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class ConnectomeDataLoader:
    """Batches of subjects, same contract as reference ``graph.py:174-197``: ``len`` is
    ``ceil(n / batch_size)``, each epoch draws ``torch.randperm`` from the global CPU generator
    (so batch membership matches the reference under the same seed) and the last batch may be
    ragged.  The dataset is packed into a :class:`SubjectStore` on first use.

    Under data parallelism (``world_size > 1``) every rank draws the same permutation and takes
    its contiguous slice of every global chunk, so the union over ranks is exactly the
    single-process batch."""

    def __init__(self, dataset, batch_size: int = 16, shuffle: bool = True, *, rank: Optional[int] = None,
                 world_size: Optional[int] = None, device=None, prepare_for: Optional[str] = None):
        self.dataset = dataset
        self.prepare_for = prepare_for   # "gcn" / "sage": collate also emits that family's aggregation structure
        self.batch_size = batch_size
        self.shuffle = shuffle
        self.device = device
        self._rank, self._world = rank, world_size
        self._store: Optional[SubjectStore] = dataset if isinstance(dataset, SubjectStore) else None
        self._store_key = None

    # -- host logic (no GPU needed) -------------------------------------------------------
    def __len__(self) -> int:
        return math.ceil(len(self.dataset) / self.batch_size)

    def _dist(self) -> tuple[int, int]:
        if self._rank is not None and self._world is not None:
            return self._rank, self._world
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
        return 0, 1

    def epoch_order(self) -> list[int]:
        n = len(self.dataset)
        return torch.randperm(n).tolist() if self.shuffle else list(range(n))

    def plan(self, order: Sequence[int]) -> list[dict]:
        """Split an epoch order into global chunks and this rank's slice of each."""
        rank, world = self._dist()
        node_ptr = self._store.node_ptr_host if self._store is not None else None
        steps = []
        for start in range(0, len(order), self.batch_size):
            chunk = np.asarray(order[start:start + self.batch_size], dtype=np.int64)
            lo, hi = shard_bounds(len(chunk), rank, world)
            step = dict(ids=chunk[lo:hi], graph_base=lo, global_num_graphs=len(chunk))
            if node_ptr is not None:
                sizes = node_ptr[chunk + 1] - node_ptr[chunk]
                step["row_base"] = int(sizes[:lo].sum())
                step["global_num_nodes"] = int(sizes.sum())
            steps.append(step)
        return steps

    # -- device side ----------------------------------------------------------------------
    def store(self) -> SubjectStore:
        if isinstance(self.dataset, SubjectStore):
            return self.dataset
        key = (id(self.dataset), len(self.dataset))
        if self._store is None or self._store_key != key:
            self._store = SubjectStore.from_graphs(self.dataset, self.device)
            self._store_key = key
        return self._store

    def invalidate(self) -> None:
        """Forget the packed copy (call after mutating ``dataset`` in place)."""
        if not isinstance(self.dataset, SubjectStore):
            self._store = None

    def __iter__(self):
        order = self.epoch_order()
        store = self.store()
        for step in self.plan(order):
            yield store.collate(step["ids"], prepare_for=self.prepare_for, row_base=step.get("row_base", 0), graph_base=step["graph_base"],
                                global_num_graphs=step["global_num_graphs"],
                                global_num_nodes=step.get("global_num_nodes"))
