"""TEST-ONLY restatement of the reference's real-data recipe (reference README.md:155-179) in plain torch on the CPU,
with `A_thresh[src, dst]` for the README's `A_thresh[src]` (see connectome_gnn/ingest.py)."""
import numpy as np
import torch


def recipe(connectivity_matrix: np.ndarray, q: float = 0.90):
    A = torch.tensor(connectivity_matrix, dtype=torch.float32)
    threshold = A.flatten().quantile(q)
    A_thresh = (A > threshold).float() * A
    src, dst = torch.where(A_thresh > 0)
    edge_index = torch.stack([torch.cat([src, dst]), torch.cat([dst, src])])
    weights = A_thresh[src, dst]
    edge_weight = torch.cat([weights, weights])
    deg = A_thresh.sum(dim=1, keepdim=True)
    node_features = deg / (deg.max() + 1e-8)
    return node_features, edge_index, edge_weight, float(threshold)


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
