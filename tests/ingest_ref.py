"""TEST-ONLY restatement of the reference's real-data recipe (reference README.md:155-179) in plain torch on the CPU,
with `A_thresh[src, dst]` for the README's `A_thresh[src]` (see connectome_gnn/ingest.py)."""
import numpy as np
import torch


def recipe(connectivity_matrix: np.ndarray, q: float = 0.90):
    """(node_features [N, 1], edge_index [2, 2 nnz], edge_weight [2 nnz], threshold) as the README recipe defines them:
    entries above the matrix's own q-quantile survive; every survivor (i, j), taken in row-major order, is listed once as
    i -> j in the first half of the edge list and once as j -> i in the second half, both with weight A[i, j]; the feature
    is the surviving row sum divided by the largest one (+1e-8)."""
    dense = torch.as_tensor(connectivity_matrix, dtype=torch.float32)
    cut = torch.quantile(dense.reshape(-1), q)
    kept = torch.where(dense > cut, dense, torch.zeros_like(dense))
    rows, cols = torch.nonzero(kept > 0, as_tuple=True)          # row-major
    w = kept[rows, cols]
    edge_index = torch.stack([torch.cat([rows, cols]), torch.cat([cols, rows])])
    edge_weight = torch.cat([w, w])
    strength = kept.sum(dim=1, keepdim=True)
    return strength / (strength.max() + 1e-8), edge_index, edge_weight, float(cut)


def random_matrices(S: int, N: int, seed: int, kind: str = "dense") -> np.ndarray:
    rng = np.random.default_rng(seed)
    out = []
    for s in range(S):
        if kind == "dense":          # symmetric, positive, zero diagonal: a tractography count matrix
            a = rng.gamma(2.0, 1.0, (N, N)).astype(np.float32)
            a = np.triu(a, 1)
            a = a + a.T
        elif kind == "sparse":       # more than 90 % zeros: the quantile is 0 and every positive entry survives
            a = rng.random((N, N)).astype(np.float32)
            a = np.where(rng.random((N, N)) < 0.04, a, 0).astype(np.float32)
            a = np.triu(a, 1)
            a = a + a.T
        elif kind == "signed":       # correlations: negative entries, asymmetric noise
            a = rng.normal(0, 1, (N, N)).astype(np.float32)
        elif kind == "ties":         # few distinct values: order statistics with ties at the rank
            a = rng.integers(0, 6, (N, N)).astype(np.float32)
        out.append(a)
    return np.stack(out)
