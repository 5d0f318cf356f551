"""The fast generator (connectome_gnn/synthetic_fast.py) against the reference generator: exact structural invariants,
statistical parity of everything the models see.  CPU only."""
import numpy as np
import pytest
import torch

from connectome_gnn.synthetic import generate_dataset
from connectome_gnn.synthetic_fast import generate_dataset_fast, generate_packed

N, S = 84, 400


@pytest.fixture(scope="module")
def both():
    return generate_dataset(num_subjects=S, num_regions=N, seed=7), generate_dataset_fast(num_subjects=S, num_regions=N, seed=7)


def test_structure_is_exact(both):
    _, fast = both
    for g in fast[:50]:
        src, dst = g.edge_index
        assert g.num_edges == 8 * N and g.node_features.shape == (N, 5)
        assert bool((src != dst).all()), "self loop"
        assert torch.equal(src[0::2], dst[1::2]) and torch.equal(dst[0::2], src[1::2]), "pairs are not adjacent reversed edges"
        assert torch.equal(g.edge_weight[0::2], g.edge_weight[1::2])
        key = torch.minimum(src, dst)[0::2] * N + torch.maximum(src, dst)[0::2]
        assert key.unique().numel() == 4 * N, "duplicate undirected pair"
        assert float(g.edge_weight.min()) > 0.0 and float(g.edge_weight.max()) < 1.0
        assert int(g.label) in (0, 1)


def test_same_seed_same_subjects():
    a = generate_packed(32, 50, seed=3)
    b = generate_packed(32, 50, seed=3)
    assert all(torch.equal(a[k], b[k]) for k in ("x", "src", "w", "label"))
    c = generate_packed(32, 50, seed=4)
    assert not torch.equal(a["w"], c["w"])


def test_statistics_match_the_reference_generator(both):
    ref, fast = both
    deg = lambda gs: np.concatenate([np.bincount(g.edge_index[1].numpy(), minlength=N) for g in gs])
    dr, df = deg(ref), deg(fast)
    assert dr.mean() == df.mean() == 8.0
    hr, hf = np.bincount(dr, minlength=20)[:20] / dr.size, np.bincount(df, minlength=20)[:20] / df.size
    assert np.abs(hr - hf).max() < 0.01, (hr, hf)                     # in-degree histogram (33 600 nodes each)
    assert abs(dr.std() - df.std()) < 0.03
    # rewired fraction: pairs that are not ring neighbours within k/2 steps
    def rewired(gs):
        out = []
        for g in gs:
            s, d = g.edge_index[:, 0::2].numpy()
            gap = np.minimum((s - d) % N, (d - s) % N)
            out.append((gap > 4).mean())
        return float(np.mean(out))
    assert abs(rewired(ref) - rewired(fast)) < 0.006, (rewired(ref), rewired(fast))
    wr, wf = torch.cat([g.edge_weight for g in ref]).double(), torch.cat([g.edge_weight for g in fast]).double()
    assert abs(float(wr.mean()) - 2 / 7) < 2e-3 and abs(float(wf.mean()) - 2 / 7) < 2e-3        # Beta(2, 5): mean 2/7
    assert abs(float(wr.var()) - float(wf.var())) < 5e-4
    xr, xf = torch.cat([g.node_features for g in ref]).double(), torch.cat([g.node_features for g in fast]).double()
    for c in range(5):
        assert abs(float(xr[:, c].mean()) - float(xf[:, c].mean())) < 0.02, c
        assert abs(float(xr[:, c].std()) - float(xf[:, c].std())) < 0.02, c
    lr, lf = np.mean([int(g.label) for g in ref]), np.mean([int(g.label) for g in fast])
    assert abs(lr - lf) < 0.12, (lr, lf)                              # two binomial(400) draws of the same p


def test_packed_forms_agree():
    """pair / compact / plain packing of the same draw describe the same edges."""
    a = generate_packed(8, 30, seed=5, compact=True, pairs=True)
    b = generate_packed(8, 30, seed=5, compact=True, pairs=False)
    c = generate_packed(8, 30, seed=5, compact=False, pairs=False)
    assert a["edge_pairs"] == 1 and b["edge_pairs"] == 0
    wb = b["src"].to(torch.int64) & 0xFFFFFFFF
    assert torch.equal(wb & 0xFFFF, c["src"].to(torch.int64)) and torch.equal(wb >> 16, c["dst"].to(torch.int64))
    wa = a["src"].to(torch.int64) & 0xFFFFFFFF
    assert torch.equal(wa & 0xFFFF, c["src"].to(torch.int64)[0::2]) and torch.equal(wa >> 16, c["dst"].to(torch.int64)[0::2])
    assert torch.equal(a["w"], c["w"][0::2]) and torch.equal(a["x"], c["x"])
