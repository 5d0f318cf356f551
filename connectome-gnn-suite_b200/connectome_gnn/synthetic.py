"""Synthetic Watts-Strogatz connectome subjects (input generator for the hot path).

Host-side mirror of the reference generator, reference ``connectome_gnn/synthetic.py:97-339``.
The message-passing kernels consume what this module produces, so it has to travel
to the GPU box (where the reference tree does not exist) and has to emit *the same
subjects* as the reference for the same seed: parity fixtures under ``tests/golden``
are generated from the reference and compared against this module bit for bit.

What is kept identical to the reference (because it is observable in the output):

* the order of draws from ``numpy.random.default_rng(seed)``: optional subject id,
  one ``random()`` per lattice edge, one ``choice`` per rewired edge, one ``beta(2, 5)``
  per undirected edge, ``lognormal``/``normal``/``normal`` node covariates, one
  ``normal(0, 2)`` for the label noise (reference ``synthetic.py:120-129,137,170-176,216``);
* the insertion / removal history of the two Python ``set`` objects that hold the
  undirected edges, because CPython's set iteration order (and therefore the COO
  edge order) is a function of that history (reference ``synthetic.py:109-130,136``).

What is different: the reference rebuilds the neighbourhood of ``u`` with a scan over
every edge for each rewiring event (``synthetic.py:124``, quadratic in E); here an
incrementally maintained adjacency table answers the same question in O(deg), which
makes 360-node subjects ~10x cheaper to draw.
"""

from __future__ import annotations

from typing import Optional

import numpy as np
import torch

__all__ = [
    "REGION_NAMES",
    "NUM_REGIONS",
    "generate_connectome",
    "generate_dataset",
    "small_world_stats",
]


# ---------------------------------------------------------------------------
# Desikan-Killiany style parcellation, 84 regions (reference synthetic.py:37-88)
# ---------------------------------------------------------------------------

def _atlas() -> list[str]:
    cortical = (
        # frontal
        "superiorfrontal rostralmiddlefrontal caudalmiddlefrontal parsopercularis "
        "parsorbitalis parstriangularis lateralorbitofrontal medialorbitofrontal precentral "
        # parietal
        "superiorparietal inferiorparietal supramarginal postcentral precuneus "
        "posteriorcingulate isthmuscingulate "
        # temporal
        "superiortemporal middletemporal inferiortemporal fusiform entorhinal "
        "parahippocampal transversetemporal "
        # occipital
        "lateraloccipital lingual cuneus pericalcarine "
        # cingulate / limbic
        "rostralanteriorcingulate caudalanteriorcingulate paracingulate"
    ).split()
    subcortical = "Thalamus Caudate Putamen Pallidum Hippocampus Amygdala Accumbens-area".split()
    tracts = "UncF ILF CST".split()

    names: list[str] = []
    for area in cortical:
        names += [f"ctx-lh-{area}", f"ctx-rh-{area}"]
    for nucleus in subcortical:
        names += [f"Left-{nucleus}", f"Right-{nucleus}"]
    names.append("Brain-Stem")
    names += ["CC_anterior", "CC_posterior"]
    for tract in tracts:
        names += [f"{tract}_left", f"{tract}_right"]
    return names


REGION_NAMES: list[str] = _atlas()
NUM_REGIONS: int = len(REGION_NAMES)  # 84

_TRAIT_NAMES = (
    "fluid_intelligence",
    "sustained_attention",
    "working_memory",
    "processing_speed",
    "cognitive_flexibility",
)


# ---------------------------------------------------------------------------
# Small-world wiring
# ---------------------------------------------------------------------------

def _small_world_pairs(n: int, k: int, beta: float, rng: np.random.Generator) -> set:
    """Undirected Watts-Strogatz pairs ``(lo, hi)``; reference ``synthetic.py:97-130``."""
    half = k // 2
    lattice: set = set()
    for i in range(n):
        for step in range(1, half + 1):
            j = (i + step) % n
            lattice.add((i, j) if i < j else (j, i))

    # `wired` starts as a copy and is edited while `lattice` is walked in set order.
    wired = set(lattice)
    nbrs: list[set] = [set() for _ in range(n)]
    for a, b in lattice:
        nbrs[a].add(b)
        nbrs[b].add(a)

    for u, v in lattice:
        if not (rng.random() < beta):
            continue
        wired.discard((u, v))
        nbrs[u].discard(v)
        nbrs[v].discard(u)
        taken = nbrs[u]
        # ascending list of nodes that are neither u nor currently adjacent to u
        free = [t for t in range(n) if t != u and t not in taken]
        if free:
            t = rng.choice(free)
            wired.add((u, t) if u < t else (t, u))
            t = int(t)
            nbrs[u].add(t)
            nbrs[t].add(u)
        else:
            wired.add((u, v))
            nbrs[u].add(v)
            nbrs[v].add(u)
    return wired


def _weighted_coo(pairs: set, rng: np.random.Generator) -> tuple[torch.Tensor, torch.Tensor]:
    """Both directions of every pair, one Beta(2,5) weight per pair (ref ``synthetic.py:133-143``)."""
    m = len(pairs)
    ends = np.empty((m, 2), dtype=np.int64)
    wts = np.empty(m, dtype=np.float64)
    for row, (a, b) in enumerate(pairs):
        ends[row, 0] = a
        ends[row, 1] = b
        wts[row] = rng.beta(2, 5)
    src = np.stack([ends[:, 0], ends[:, 1]], axis=1).reshape(-1)
    dst = np.stack([ends[:, 1], ends[:, 0]], axis=1).reshape(-1)
    edge_index = torch.from_numpy(np.stack([src, dst], axis=0))
    edge_weight = torch.from_numpy(np.repeat(wts, 2)).to(torch.float32)
    return edge_index, edge_weight


def _zscore(v: torch.Tensor) -> torch.Tensor:
    return (v - v.mean()) / (v.std() + 1e-8)


def _region_covariates(n: int, edge_index: torch.Tensor, edge_weight: torch.Tensor,
                       rng: np.random.Generator) -> torch.Tensor:
    """Five per-region features (ref ``synthetic.py:150-183``): strength/max, mean
    incident weight, z-scored volume proxy, activation proxy, z-scored thickness proxy."""
    rows = edge_index[0]
    strength = torch.zeros(n).scatter_add_(0, rows, edge_weight)
    volume = torch.tensor(rng.lognormal(mean=7.5, sigma=0.5, size=n), dtype=torch.float32)
    activation = torch.tensor(rng.normal(0, 1, size=n), dtype=torch.float32)
    thickness = torch.tensor(rng.normal(2.5, 0.3, size=n).clip(1.5, 4.0), dtype=torch.float32)
    incident = torch.zeros(n).scatter_add_(0, rows, torch.ones(rows.shape[0]))
    columns = [
        strength / (strength.max() + 1e-8),
        strength / (incident + 1e-8),
        _zscore(volume),
        activation,
        _zscore(thickness),
    ]
    return torch.stack(columns, dim=1)


def _trait_label(features: torch.Tensor, edge_weight: torch.Tensor, trait_idx: int,
                 rng: np.random.Generator) -> torch.Tensor:
    """Noisy linear trait score thresholded at zero (ref ``synthetic.py:193-218``)."""
    summary = (
        features[:, 0].mean().item(),
        edge_weight.mean().item(),
        features[:, 1].mean().item(),
    )
    loading = np.random.default_rng(trait_idx * 1337).normal(0, 1, 3)
    score = loading[0] * summary[0] + loading[1] * summary[1] + loading[2] * summary[2]
    score += rng.normal(0, 2.0)
    return torch.tensor(int(score > 0), dtype=torch.long)


# ---------------------------------------------------------------------------
# Public API (same signatures as the reference, synthetic.py:225-301)
# ---------------------------------------------------------------------------

def generate_connectome(
    num_regions: int = NUM_REGIONS,
    k: int = 8,
    beta: float = 0.15,
    trait_idx: int = 0,
    subject_id: Optional[str] = None,
    seed: Optional[int] = None,
):
    """One synthetic subject as a :class:`~connectome_gnn.graph.ConnectomeGraph`."""
    from connectome_gnn.graph import ConnectomeGraph

    rng = np.random.default_rng(seed)
    if subject_id is None:
        subject_id = f"sub-{rng.integers(10000, 99999)}"
    pairs = _small_world_pairs(num_regions, k, beta, rng)
    edge_index, edge_weight = _weighted_coo(pairs, rng)
    features = _region_covariates(num_regions, edge_index, edge_weight, rng)
    label = _trait_label(features, edge_weight, trait_idx, rng)
    return ConnectomeGraph(
        node_features=features,
        edge_index=edge_index,
        edge_weight=edge_weight,
        label=label,
        subject_id=subject_id,
    )


def generate_dataset(
    num_subjects: int = 200,
    num_regions: int = NUM_REGIONS,
    k: int = 8,
    beta: float = 0.15,
    trait_idx: int = 0,
    seed: int = 42,
) -> list:
    """``num_subjects`` subjects; per-subject seeds are drawn from the master seed."""
    master = np.random.default_rng(seed)
    subject_seeds = master.integers(0, 2**31, size=num_subjects).tolist()
    out = []
    for i, s in enumerate(subject_seeds):
        out.append(
            generate_connectome(
                num_regions=num_regions, k=k, beta=beta, trait_idx=trait_idx,
                subject_id=f"sub-{i:04d}", seed=int(s),
            )
        )
    return out


def small_world_stats(graphs: list) -> dict:
    """Mean weighted clustering proxy and mean BFS path length from the first 20 regions
    of each subject (ref ``synthetic.py:304-339``). Diagnostic only, not on the hot path."""
    clustering, path_len = [], []
    for g in graphs:
        A = g.adjacency_matrix().numpy()
        n = A.shape[0]
        strength = A.sum(axis=1)
        closed = np.diagonal(A @ A @ A)
        denom = strength * (strength - 1)
        with np.errstate(divide="ignore", invalid="ignore"):
            coeff = np.where(denom > 0, closed / denom, 0.0)
        clustering.append(float(coeff.mean()))

        linked = A > 0
        hops_total, hops_count = 0, 0
        for root in range(min(20, n)):
            seen = np.zeros(n, dtype=bool)
            seen[root] = True
            frontier = seen.copy()
            depth = 0
            while frontier.any():
                depth += 1
                reach = linked[frontier].any(axis=0) & ~seen
                cnt = int(reach.sum())
                hops_total += depth * cnt
                hops_count += cnt
                seen |= reach
                frontier = reach
        path_len.append(hops_total / hops_count if hops_count else float("nan"))
    return {
        "mean_clustering": float(np.mean(clustering)),
        "mean_avg_path_length": float(np.nanmean(path_len)),
        "num_graphs": len(graphs),
    }
