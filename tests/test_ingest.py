"""Dense-matrix ingestion (csrc/ingest.cu) against the reference's README recipe restated in torch (tests/ingest_ref.py):
edge list (order, duplicates), weights and threshold bit-exact, the degree feature to fp32 summation order."""
import numpy as np
import pytest
import torch

import helpers
import ingest_ref


def _check(device, S, N, kind, seed=0):
    from connectome_gnn.ingest import matrices_to_packed
    mats = ingest_ref.random_matrices(S, N, seed, kind)
    p = matrices_to_packed(mats, labels=list(range(S)), device=device)
    assert p["node_ptr"].tolist() == [N * s for s in range(S + 1)] and p["num_features"] == 1
    for s in range(S):
        x, ei, ew, thr = ingest_ref.recipe(mats[s])
        e0, e1 = int(p["edge_ptr"][s]), int(p["edge_ptr"][s + 1])
        assert e1 - e0 == ei.shape[1], (kind, s, e1 - e0, ei.shape[1])
        assert float(p["threshold"][s]) == thr, (kind, s)
        assert torch.equal(p["src"][e0:e1].cpu().long(), ei[0]) and torch.equal(p["dst"][e0:e1].cpu().long(), ei[1]), (kind, s)
        assert torch.equal(p["w"][e0:e1].cpu(), ew), (kind, s)
        helpers.assert_close(p["x"][s * N:(s + 1) * N].cpu(), x, f"{kind} subject {s}: degree feature", tol=1e-6)
    return p


@pytest.mark.parametrize("kind", ["dense", "sparse", "signed", "ties"])
def test_ingest_matches_the_recipe_on_the_simulator(on_emu, kind):
    _check("cpu", 3, 20, kind)


def test_ingested_subjects_collate_and_run_on_the_simulator(on_emu):
    """The packed result is a SubjectStore like any other: duplicated edges and all (collate vs the host collate of the
    same graphs, bit for bit)."""
    from connectome_gnn.graph import ConnectomeGraph, SubjectStore, collate_graphs
    p = _check("cpu", 3, 20, "dense", seed=4)
    store = SubjectStore(p, "cpu")
    b = store.collate(np.array([2, 0, 1]))
    graphs = []
    for s in (2, 0, 1):
        x, ei, ew, _ = ingest_ref.recipe(ingest_ref.random_matrices(3, 20, 4, "dense")[s])
        graphs.append(ConnectomeGraph(x, ei, ew, torch.tensor(s)))
    ref = collate_graphs(graphs)
    helpers.assert_close(b.node_features, ref.node_features, "features", tol=1e-6)
    for f in ("edge_index", "edge_weight", "batch", "labels", "ptr"):
        assert torch.equal(getattr(b, f), getattr(ref, f)), f


@pytest.mark.gpu
@pytest.mark.parametrize("N,kind", [(84, "dense"), (360, "dense"), (360, "sparse"), (84, "signed"), (84, "ties"), (33, "dense")])
def test_ingest_matches_the_recipe(N, kind):
    _check("cuda", 4, N, kind)


@pytest.mark.gpu
def test_ingested_connectomes_train():
    """360 x 360 matrices -> store -> GCN / GraphSAGE steps (dense-input and duplicate-edge case of SURVEY 8f rank 4)."""
    from connectome_gnn.graph import SubjectStore
    from connectome_gnn.ingest import matrices_to_packed, hcp_matrix_to_graph
    from connectome_gnn.models import GCNConnectome, GraphSAGEConnectome
    from connectome_gnn.train import Trainer
    mats = ingest_ref.random_matrices(8, 360, 9, "dense")
    store = SubjectStore(matrices_to_packed(mats, labels=[0, 1] * 4, device="cuda"), "cuda")
    for cls in (GCNConnectome, GraphSAGEConnectome):
        torch.manual_seed(0)
        m = cls(in_channels=1, hidden_dim=64, num_classes=2, num_layers=3, dropout=0.1).to("cuda")
        tr = Trainer(m, torch.optim.Adam(m.parameters(), lr=1e-2), device="cuda")
        m.train()
        losses = [float(tr.train_step(store.collate(np.arange(8), prepare_for=m.kind))) for _ in range(3)]
        assert all(np.isfinite(losses)) and losses[-1] < losses[0] + 1e-3, losses
    g = hcp_matrix_to_graph(mats[0], 1)
    x, ei, ew, _ = ingest_ref.recipe(mats[0])
    assert torch.equal(g.edge_index, ei) and torch.equal(g.edge_weight, ew) and int(g.label) == 1
