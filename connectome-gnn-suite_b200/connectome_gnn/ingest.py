"""Real-connectome ingestion: dense connectivity matrices -> subjects.

The reference documents, besides its generator, one recipe for real data (reference ``README.md:145-179``,
``hcp_matrix_to_graph``): threshold a dense 84 x 84 / 360 x 360 structural connectivity matrix at its own 90th
percentile, list every surviving entry in both directions, use the normalised weighted degree as the node feature.
Here the thresholding, the stable row-major compaction and the feature run on the device, one CTA per subject
(``csrc/ingest.cu``: ``cgnn_ingest_threshold`` + ``cgnn_ingest_emit``), for any number of subjects at once; the result is
the packed arena ``SubjectStore`` takes - no per-subject Python object is built - or, through ``hcp_matrix_to_graph``, the
``ConnectomeGraph`` the README's function returns.

One deliberate correction: the README computes the weights as ``A_thresh[src]`` - a row lookup that yields an
``[nnz, N]`` matrix, which ``ConnectomeGraph`` cannot hold; the evident intent ``A_thresh[src, dst]`` is what is built.
Edge order, duplicates (a symmetric matrix lists every undirected edge four times) and weights are bit-exact against
that recipe (``tests/test_ingest.py``).
"""

from __future__ import annotations

from typing import Optional, Sequence, Union

import numpy as np
import torch

from . import _engine
from .graph import ConnectomeGraph

__all__ = ["matrices_to_packed", "hcp_matrix_to_graph"]


def matrices_to_packed(matrices, labels: Optional[Sequence[int]] = None, quantile: float = 0.90, device=None) -> dict:
    """``matrices`` [S, N, N] (numpy or torch, any device) -> the packed arena of ``SubjectStore`` (``graph.pack_graphs``'s
    dict; ``x`` / ``src`` / ``dst`` / ``w`` stay on the device they were built on, the pointer tables are host tensors).
    Subject s keeps the entries of its matrix above that matrix's ``quantile`` (``torch.quantile`` semantics)."""
    if isinstance(matrices, np.ndarray):
        matrices = torch.from_numpy(np.ascontiguousarray(matrices, dtype=np.float32))
    if matrices.dim() == 2:
        matrices = matrices.unsqueeze(0)
    if matrices.dim() != 3 or matrices.shape[1] != matrices.shape[2]:
        raise ValueError("expected [S, N, N] connectivity matrices")
    dev = torch.device(device) if device is not None else (matrices.device if matrices.is_cuda else _engine.default_device())
    A = matrices.to(device=dev, dtype=torch.float32).contiguous()
    S, N = int(A.shape[0]), int(A.shape[1])
    if labels is not None and len(labels) != S:
        raise ValueError("one label per matrix")
    eng = _engine.engine_for(A)
    thr = torch.empty(S, dtype=torch.float32, device=dev)
    sel = torch.empty(S, dtype=torch.int32, device=dev)
    eng._call("cgnn_ingest_threshold", A.data_ptr(), S, N, float(quantile), thr.data_ptr(), sel.data_ptr(), eng.stream())
    edge_ptr = torch.zeros(S + 1, dtype=torch.int64)
    edge_ptr[1:] = torch.cumsum(sel.cpu().to(torch.int64) * 2, 0)          # the one host read: sizes of the outputs
    E = int(edge_ptr[-1])
    src = torch.empty(max(E, 1), dtype=torch.int32, device=dev)      # (never a null pointer: a matrix may keep nothing)
    dst = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
    w = torch.empty(max(E, 1), dtype=torch.float32, device=dev)
    x = torch.empty((S * N, 1), dtype=torch.float32, device=dev)
    eptr_dev = edge_ptr.to(dev)
    eng._call("cgnn_ingest_emit", A.data_ptr(), S, N, thr.data_ptr(), eptr_dev.data_ptr(), src.data_ptr(), dst.data_ptr(),
              w.data_ptr(), x.data_ptr(), eng.stream())
    label = torch.zeros(S, dtype=torch.int64)
    has_label = np.zeros(S, dtype=bool)
    if labels is not None:
        label = torch.as_tensor(np.asarray(labels), dtype=torch.int64)
        has_label[:] = True
    src, dst, w = src[:E], dst[:E], w[:E]
    return dict(x=x, src=src, dst=dst, w=w, node_ptr=torch.arange(S + 1, dtype=torch.int64) * N, edge_ptr=edge_ptr, label=label,
                has_label=has_label, num_features=1, edge_pairs=0, threshold=thr)


def hcp_matrix_to_graph(connectivity_matrix: Union[np.ndarray, torch.Tensor], label: int) -> ConnectomeGraph:
    """The README's function (``README.md:155-179``), computed on the device; returns host tensors like the recipe."""
    p = matrices_to_packed(connectivity_matrix, [label])
    ei = torch.stack([p["src"], p["dst"]]).to(torch.int64).cpu()
    return ConnectomeGraph(node_features=p["x"].cpu(), edge_index=ei, edge_weight=p["w"].cpu(),
                           label=torch.tensor(int(label), dtype=torch.long))
