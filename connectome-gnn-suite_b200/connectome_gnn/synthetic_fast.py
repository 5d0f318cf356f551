"""Fast synthetic connectomes for the scale configurations.

``connectome_gnn.synthetic`` restates the reference generator (reference ``synthetic.py:97-263``) subject by subject
and bit for bit; at 64 ms per 360-node subject that is 18 hours for the million subjects of BASELINE configs[3].  This
module draws the SAME distribution - Watts-Strogatz rewiring of a k-ring, one Beta(2, 5) weight per undirected pair,
the five regional covariates, the noisy linear label - for thousands of subjects at a time with tensor operations (CPU
or CUDA; torch is only the array library here, nothing on the message-passing path), and hands back the packed arena
``SubjectStore`` takes, so no per-subject Python object is ever built.

Parity is statistical, not bitwise (the reference's edge order comes from Python ``set`` iteration and its stream from
numpy's PCG64): ``tests/test_synthetic_fast.py`` compares degree histogram, weight / feature moments, clustering and label
balance with the reference generator, and checks the structural invariants exactly (8 N directed edges per subject,
adjacent reversed pairs of equal weight, no self loops, no duplicates).

One deliberate difference in mechanism: the reference rewires a subject's pairs one after the other, each seeing the
edges placed before it; here all pairs chosen for rewiring (probability ``beta`` each) are lifted at once and re-placed by
rejection sampling against the remaining graph and against each other.  Both produce a uniformly chosen free target per
rewired pair; the difference is second order in ``beta * k / N``.
"""

from __future__ import annotations

from typing import Optional

import numpy as np
import torch

__all__ = ["generate_packed", "generate_dataset_fast"]


def _chunk(n_subjects: int, n: int, k: int, beta: float, loading: np.ndarray, gen: torch.Generator, device) -> dict:
    """One chunk of subjects: returns x [S, n, 5] f32, u / v [S, m] int64 (one row per undirected pair), w [S, m] f32, label [S]."""
    S, half = n_subjects, k // 2
    m = n * half
    base_u = torch.arange(n, device=device).repeat_interleave(half)
    steps = torch.arange(1, half + 1, device=device).repeat(n)
    base_v = (base_u + steps) % n
    u = base_u.expand(S, m).contiguous()
    v = base_v.expand(S, m).contiguous()
    sidx = torch.arange(S, device=device).unsqueeze(1).expand(S, m)
    adj = torch.zeros(S, n, n, dtype=torch.bool, device=device)
    adj[sidx, u, v] = True
    adj[sidx, v, u] = True
    rew = torch.rand(S, m, generator=gen, device=device) < beta
    rs, re = rew.nonzero(as_tuple=True)
    # lift every pair chosen for rewiring, then place each on a uniformly drawn free partner of its first endpoint
    adj[rs, u[rs, re], v[rs, re]] = False
    adj[rs, v[rs, re], u[rs, re]] = False
    for _ in range(200):
        if rs.numel() == 0:
            break
        uu = u[rs, re]
        t = torch.randint(0, n, (rs.numel(),), generator=gen, device=device)
        ok = (t != uu) & ~adj[rs, uu, t]
        # two pairs of one subject must not land on the same undirected pair in this round: keep the first of each key
        lo, hi = torch.minimum(uu, t), torch.maximum(uu, t)
        key = (rs * n + lo) * n + hi
        key = torch.where(ok, key, -1 - torch.arange(rs.numel(), device=device))     # rejected draws never collide
        order = torch.argsort(key, stable=True)
        sk = key[order]
        first = torch.ones_like(sk, dtype=torch.bool)
        first[1:] = sk[1:] != sk[:-1]
        keep = torch.zeros_like(ok)
        keep[order] = first
        ok &= keep
        a_s, a_e, a_u, a_t = rs[ok], re[ok], uu[ok], t[ok]
        adj[a_s, a_u, a_t] = True
        adj[a_s, a_t, a_u] = True
        v[a_s, a_e] = a_t
        rs, re = rs[~ok], re[~ok]
    if rs.numel():      # no free partner found (tiny graphs): the pair goes back where it was, as in the reference
        adj[rs, u[rs, re], v[rs, re]] = True
        adj[rs, v[rs, re], u[rs, re]] = True
    del adj
    # one Beta(2, 5) weight per pair: Beta(a, b) = G_a / (G_a + G_b) with unit-scale gammas
    ga = torch._standard_gamma(torch.full((S, m), 2.0, dtype=torch.float64, device=device), generator=gen)
    gb = torch._standard_gamma(torch.full((S, m), 5.0, dtype=torch.float64, device=device), generator=gen)
    w = (ga / (ga + gb)).to(torch.float32)
    # the five regional covariates (reference synthetic.py:150-183)
    strength = torch.zeros(S, n, device=device).scatter_add_(1, u, w).scatter_add_(1, v, w)
    incident = torch.zeros(S, n, device=device).scatter_add_(1, u, torch.ones_like(w)).scatter_add_(1, v, torch.ones_like(w))
    normal = lambda mean, std: (torch.randn(S, n, generator=gen, device=device, dtype=torch.float64) * std + mean)
    volume = torch.exp(normal(7.5, 0.5)).to(torch.float32)
    activation = normal(0.0, 1.0).to(torch.float32)
    thickness = normal(2.5, 0.3).clamp(1.5, 4.0).to(torch.float32)
    zscore = lambda t_: (t_ - t_.mean(dim=1, keepdim=True)) / (t_.std(dim=1, keepdim=True) + 1e-8)
    x = torch.stack([strength / (strength.max(dim=1, keepdim=True).values + 1e-8), strength / (incident + 1e-8),
                     zscore(volume), activation, zscore(thickness)], dim=2)
    # noisy linear trait, thresholded at zero (reference synthetic.py:193-218); every pair counts twice in the directed mean
    score = (loading[0] * x[:, :, 0].mean(dim=1).double() + loading[1] * w.mean(dim=1).double() +
             loading[2] * x[:, :, 1].mean(dim=1).double())
    score = score + torch.randn(S, generator=gen, device=device, dtype=torch.float64) * 2.0
    return dict(x=x, u=u, v=v, w=w, label=(score > 0).to(torch.int64))


def generate_packed(num_subjects: int, num_regions: int = 84, k: int = 8, beta: float = 0.15, trait_idx: int = 0,
                    seed: int = 42, *, device="cpu", compact: bool = True, pairs: bool = True, chunk: Optional[int] = None) -> dict:
    """``num_subjects`` synthetic subjects as the packed arena ``SubjectStore`` takes (the dict ``graph.pack_graphs``
    returns, CPU tensors): ``compact`` packs both endpoints of an edge into one int32, ``pairs`` additionally stores one
    entry per undirected pair - the generator emits every pair as two adjacent reversed edges, so that form is lossless.
    ``device`` is where the arithmetic runs ("cuda": a million 84-node subjects in seconds)."""
    n, S = int(num_regions), int(num_subjects)
    if k % 2 or k < 2 or n <= k:
        raise ValueError("need an even k >= 2 and more regions than k")
    if compact and n > 65535:
        raise ValueError("compact packing needs < 65536 regions")
    device = torch.device(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    loading = np.random.default_rng(trait_idx * 1337).normal(0, 1, 3)
    m = n * (k // 2)
    if chunk is None:
        chunk = max(1, min(S, (256 << 20) // (n * n)))       # the boolean adjacency of a chunk stays under 256 MB
    xs, us, vs, ws, ys = [], [], [], [], []
    for lo in range(0, S, chunk):
        c = _chunk(min(chunk, S - lo), n, k, beta, loading, gen, device)
        xs.append(c["x"].reshape(-1, 5).cpu()); us.append(c["u"].cpu()); vs.append(c["v"].cpu())
        ws.append(c["w"].cpu()); ys.append(c["label"].cpu())
    x, u, v, w, label = torch.cat(xs), torch.cat(us), torch.cat(vs), torch.cat(ws), torch.cat(ys)
    node_ptr = torch.arange(S + 1, dtype=torch.int64) * n
    edge_ptr = torch.arange(S + 1, dtype=torch.int64) * (2 * m)
    if compact and pairs:
        word = (u | (v << 16))
        src = torch.where(word >= 2 ** 31, word - 2 ** 32, word).to(torch.int32).reshape(-1).contiguous()
        dst, wt, edge_pairs = torch.zeros(0, dtype=torch.int32), w.reshape(-1).contiguous(), 1
    else:
        s2 = torch.stack([u, v], dim=2).reshape(S, 2 * m)      # (u -> v), (v -> u) adjacent
        d2 = torch.stack([v, u], dim=2).reshape(S, 2 * m)
        wt, edge_pairs = w.repeat_interleave(2, dim=1).reshape(-1).contiguous(), 0
        if compact:
            word = (s2 | (d2 << 16)).reshape(-1)
            src = torch.where(word >= 2 ** 31, word - 2 ** 32, word).to(torch.int32).contiguous()
            dst = torch.zeros(0, dtype=torch.int32)
        else:
            src, dst = s2.reshape(-1).to(torch.int32).contiguous(), d2.reshape(-1).to(torch.int32).contiguous()
    return dict(x=x.contiguous(), src=src, dst=dst, w=wt, node_ptr=node_ptr, edge_ptr=edge_ptr, label=label,
                has_label=np.ones(S, dtype=bool), num_features=5, edge_pairs=edge_pairs)


def generate_dataset_fast(num_subjects: int = 200, num_regions: int = 84, k: int = 8, beta: float = 0.15,
                          trait_idx: int = 0, seed: int = 42, device="cpu") -> list:
    """The same subjects as a list of ``ConnectomeGraph`` (same signature as ``generate_dataset``; for moderate counts)."""
    from .graph import ConnectomeGraph
    p = generate_packed(num_subjects, num_regions, k, beta, trait_idx, seed, device=device, compact=False, pairs=False)
    n, e = num_regions, 2 * num_regions * (k // 2)
    out = []
    for i in range(num_subjects):
        ei = torch.stack([p["src"][i * e:(i + 1) * e], p["dst"][i * e:(i + 1) * e]]).to(torch.int64)
        out.append(ConnectomeGraph(p["x"][i * n:(i + 1) * n], ei, p["w"][i * e:(i + 1) * e], p["label"][i].clone(), f"sub-{i:04d}"))
    return out
