"""Kernel-logic parity on the test-only simulator (tests/emu): the SAME csrc/*.cu sources, compiled by
g++ with every CUDA thread a fiber, driven through the same C ABI and the same host package code.
Catches indexing / barrier / reduction bugs without a GPU; says nothing about speed, and the package
never takes this path (see conftest.on_emu)."""
import numpy as np
import pytest
import torch

import helpers
import parity


@pytest.mark.parametrize("fixture", ["ref_small.npz", "ref_ragged.npz", "ref_c1.npz"])
def test_collate_bit_exact(on_emu, fixture):
    parity.check_collate(helpers.golden(fixture), "cpu")


@pytest.mark.parametrize("fixture", ["ref_small.npz", "ref_ragged.npz"])
@pytest.mark.parametrize("kind", ["gcn", "sage"])
def test_model_parity(on_emu, fixture, kind):
    parity.check_model(helpers.golden(fixture), kind, "cpu")


@pytest.mark.parametrize("kind", ["gcn", "sage"])
def test_trainer_trajectory(on_emu, kind):
    parity.check_trainer(kind, "cpu")


def test_csr_from_coo_equals_collate(on_emu):
    """A batch moved from the host / assembled by hand gets the same CSR as the collate kernel builds."""
    from connectome_gnn.graph import ConnectomeBatch, collate_graphs
    a = helpers.golden("ref_ragged.npz")
    b = collate_graphs(helpers.graphs_from_store(a))
    hand = ConnectomeBatch(b.node_features, b.edge_index, b.edge_weight, b.batch, b.labels, b.ptr)
    c = hand.ensure_csr()
    for f in ("in_rowptr", "in_col", "in_w", "in_wn", "out_rowptr", "out_col", "out_w", "out_wn", "deg", "dinv", "wsum", "eptr"):
        assert torch.equal(getattr(c, f), getattr(b.csr, f)), f
    assert c.max_nodes == b.csr.max_nodes


def test_dropout_masks_regenerate_and_gradients_are_consistent(on_emu):
    """Dropout on: same seed -> same output; backward regenerates the forward mask (directional
    finite-difference check of d loss / d theta with the mask held fixed by the seed)."""
    from connectome_gnn.graph import collate_graphs
    from connectome_gnn.train import CrossEntropyLoss
    a = helpers.golden("ref_small.npz")
    b = collate_graphs(helpers.graphs_from_store(a))
    for kind in ("gcn", "sage"):
        m = parity.make_model(kind, a, "cpu", dropout=0.3)
        m.train()
        for bn in m.batch_norms:
            bn.momentum = 0.0            # keep running stats fixed across the repeated forwards

        def loss_at():
            torch.manual_seed(123)       # fixes the dropout stream seed drawn by the model
            return CrossEntropyLoss()(m(b), b.labels)

        l0 = loss_at()
        l0.backward()
        assert float(loss_at()) == float(l0), "same seed must give the same masks"
        torch.manual_seed(124)
        assert float(CrossEntropyLoss()(m(b), b.labels)) != float(l0), "different seed, different masks"
        params = [p for p in m.parameters()]
        direction = [torch.randn(p.shape, generator=torch.Generator().manual_seed(i)) for i, p in enumerate(params)]
        analytic = sum(float((p.grad * d).sum()) for p, d in zip(params, direction))
        eps = 2e-4   # small enough that ReLU kinks crossed by the probe stay below the tolerance (checked: converges to analytic)
        with torch.no_grad():
            for p, d in zip(params, direction): p.add_(eps * d)
            lp = float(loss_at())
            for p, d in zip(params, direction): p.sub_(2 * eps * d)
            lm = float(loss_at())
        numeric = (lp - lm) / (2 * eps)
        assert numeric == pytest.approx(analytic, rel=0.05, abs=2e-3), (kind, numeric, analytic)


def test_dropout_keep_rate(on_emu, emu_engine):
    """Mean of dropout(ones) over many elements ~ 1 (unbiased scaling), keep fraction ~ 1 - p."""
    from connectome_gnn._engine import Act
    n, C = 400, 64
    ones = torch.ones(n, C)
    ptr = torch.tensor([0, n], dtype=torch.int64)
    ident = (torch.ones(C), torch.zeros(C))
    for p in (0.1, 0.3, 0.5):
        emb = emu_engine.pool_fwd(ones, Act(ident[0], ident[1], False, p, seed=99, site=1), ptr, 1)
        keep = float(emb.mean()) * (1 - p)
        assert abs(keep - (1 - p)) < 0.01, (p, keep)


def test_syncbn_statistics_merge(on_emu, emu_engine):
    """Two 'ranks' (halves of a batch) -> merged statistics == statistics of the whole batch (Chan merge),
    which is what makes data-parallel BatchNorm equal the reference's single-process batch semantics."""
    from connectome_gnn._engine import Act
    from connectome_gnn.graph import SubjectStore
    a = helpers.golden("ref_small.npz")
    graphs = helpers.graphs_from_store(a)
    store = SubjectStore.from_graphs(graphs, device="cpu")
    W = torch.from_numpy(a["gcn.init.convs.0.linear.weight"]); bias = torch.zeros(W.shape[0])
    def stats_of(ids):
        b = store.collate(ids)
        _, st, _ = emu_engine.layer_fwd("gcn", b.node_features, Act(), W, bias, b.csr, b.ptr, b.num_graphs, True)
        return st
    whole = stats_of(np.arange(8))
    merged = emu_engine.bn_merge_stats(torch.stack([stats_of(np.arange(0, 5)), stats_of(np.arange(5, 8))]), W.shape[0])
    assert float(whole[0]) == float(merged[0]) == 160.0
    assert torch.allclose(whole, merged, rtol=1e-6, atol=1e-9)


def test_oracle_random_weights(on_emu):
    from connectome_gnn.synthetic import generate_dataset
    graphs = generate_dataset(num_subjects=5, num_regions=70, seed=11)   # 70 rows: chunk boundary (64) inside a subject
    parity.check_against_oracle(graphs, "gcn", "cpu", hidden=20, layers=2)  # hidden not a multiple of 32
    parity.check_against_oracle(graphs, "sage", "cpu", hidden=12, layers=2)


def test_pair_store_collates_identically(on_emu):
    """One entry per undirected edge (cgnn_store_t.edge_pairs) expands to the same batch and CSR bit for bit."""
    import numpy as np
    from connectome_gnn.graph import SubjectStore, pack_graphs
    from connectome_gnn.synthetic import generate_dataset
    graphs = generate_dataset(num_subjects=5, num_regions=20, seed=21)
    ids = np.array([3, 0, 4, 4, 1])
    packed = pack_graphs(graphs, compact=True, pairs=True)
    assert packed["edge_pairs"] == 1 and packed["src"].numel() * 2 == int(packed["edge_ptr"][-1])
    a = SubjectStore(pack_graphs(graphs), "cpu").collate(ids, prepare_for="gcn")
    b = SubjectStore(packed, "cpu").collate(ids, prepare_for="gcn")
    for f in ("node_features", "edge_index", "edge_weight", "batch", "labels", "ptr"):
        assert torch.equal(getattr(a, f), getattr(b, f)), f
    for f in ("in_rowptr", "in_col", "in_w", "in_wn", "out_rowptr", "out_col", "out_w", "out_wn", "deg", "dinv", "wsum"):
        assert torch.equal(getattr(a.csr, f), getattr(b.csr, f)), f


# ---- warp-specialised forward engine (csrc/engine.cu) on the simulator: mbarrier protocol, TMA box layout, tensor-memory
# ---- lane quadrants, unit tables and the XOR-swizzled gather, against the reference fixtures and the oracle ----------

def test_ws_engine_reference_fixture_c1(on_emu):
    """16 x 84-node subjects, hidden 64: four subjects per unit, a CTA walks several units."""
    parity.check_model(helpers.golden("ref_c1.npz"), "gcn", "cpu")


@pytest.mark.parametrize("shape", [(2, 360, 3), (5, 70, 2), (3, 130, 2)])
def test_ws_engine_oracle_shapes(on_emu, shape):
    """360-node subjects (three tiles, one subject per unit), 70-node (five per unit, short last unit) and 130-node
    (two per unit, a subject straddling a tile boundary) against oracle/port.py, hidden 64."""
    from connectome_gnn.synthetic import generate_dataset
    subjects, regions, layers = shape
    graphs = generate_dataset(num_subjects=subjects, num_regions=regions, seed=11)
    parity.check_against_oracle(graphs, "gcn", "cpu", hidden=64, layers=layers)


def test_ws_engine_mixed_sizes_and_empty_subject(on_emu):
    """Subjects of different sizes in one unit, one of them with no nodes at all."""
    from connectome_gnn.graph import ConnectomeGraph
    from connectome_gnn.synthetic import generate_connectome
    graphs = [generate_connectome(num_regions=n, seed=s) for s, n in enumerate((33, 90, 12, 57, 101, 64))]
    empty = ConnectomeGraph(torch.zeros(0, 5), torch.zeros(2, 0, dtype=torch.int64), torch.zeros(0), torch.tensor(1), "sub-empty")
    graphs.insert(2, empty)
    parity.check_against_oracle(graphs, "gcn", "cpu", hidden=64, layers=2)


def test_ws_engine_dropout_matches_generic_kernels(on_emu):
    """With dropout on, the engine (converters apply the mask thread-per-row) and the generic kernels regenerate the
    same masks from (seed, site, row, channel): identical outputs up to fp32 summation order."""
    from connectome_gnn.graph import collate_graphs
    from connectome_gnn.models import GCNConnectome
    from connectome_gnn.synthetic import generate_dataset
    b = collate_graphs(generate_dataset(num_subjects=3, num_regions=84, seed=4))
    torch.manual_seed(0)
    m = GCNConnectome(in_channels=5, hidden_dim=64, num_classes=2, num_layers=3, dropout=0.4).train()
    for bn in m.batch_norms:
        bn.momentum = 0.0
    outs = []
    for tc in (1, 0):
        assert on_emu.lib.cgnn_set_option(1, tc) == 0
        try:
            torch.manual_seed(7)
            outs.append(m(b).detach().clone())
        finally:
            on_emu.lib.cgnn_set_option(1, 1)
    assert float((outs[0] - outs[1]).abs().max()) <= 1e-5 * float(outs[1].abs().max())


@pytest.mark.parametrize("compact", [None, False])      # pair store (k_collate_pairs) / plain store (k_collate_graph)
def test_lean_collate(on_emu, compact):
    # hidden 64: the GCN hidden layers run on the warp-specialised engine, which never reads the CSR arrays
    parity.check_lean_collate("cpu", "gcn", compact)


def test_multi_subject_units(on_emu):
    parity.check_multi_subject_units("cpu", subjects=10, regions=30)


def test_pooled_last_layer(on_emu):
    parity.check_pooled_last_layer("cpu", sizes=(84, 30, 130, 57, 200, 10, 12))


def test_pair_collate_bit_exact(on_emu):
    parity.check_pair_collate("cpu")


def test_pair_collate_fuzz(on_emu):
    parity.check_pair_collate_fuzz("cpu", seed=1, rounds=3)


@pytest.mark.parametrize("kind", ["gcn", "sage"])
def test_batches_die_by_refcount(on_emu, kind):
    parity.check_batches_die_by_refcount("cpu", kind)


@pytest.mark.parametrize("layers", [2, 3, 4])
def test_fused_eval(on_emu, layers):
    """K9 on the simulator: ring / weight-reload / readout protocol for 2, 3 and 4 layers, mixed subject sizes."""
    parity.check_fused_eval("cpu", layers, sizes=(84, 30, 130, 57, 84, 200))


@pytest.mark.parametrize("kind", ["gcn", "sage"])
def test_fused_step_flat_buffers(on_emu, kind):
    parity.check_fused_step("cpu", kind, steps=3)


def test_adam_kernel(on_emu):
    parity.check_adam_kernel("cpu")


def test_trainer_asks_its_loader_for_lean_batches(on_emu):
    """Trainer.train_epoch / evaluate consume the loader's batches themselves: a ConnectomeDataLoader without an explicit
    prepare_for collates for the trainer's model family while it is being driven by the trainer, and is left as it was."""
    from connectome_gnn.graph import ConnectomeDataLoader, SubjectStore
    from connectome_gnn.models import GraphSAGEConnectome
    from connectome_gnn.synthetic import generate_dataset
    from connectome_gnn.train import Trainer
    graphs = generate_dataset(num_subjects=12, num_regions=20, seed=3)
    loader = ConnectomeDataLoader(graphs, batch_size=4, shuffle=False)
    seen = []
    orig = SubjectStore.collate

    def spy(self, ids, *a, **k):
        if "ids_device" not in k:                       # (a lean batch filling itself in calls collate again: not the loader)
            seen.append(k.get("prepare_for"))
        return orig(self, ids, *a, **k)
    SubjectStore.collate = spy
    try:
        torch.manual_seed(0)
        m = GraphSAGEConnectome(in_channels=5, hidden_dim=32, num_classes=2, num_layers=2, dropout=0.0)
        tr = Trainer(m, torch.optim.SGD(m.parameters(), lr=0.01), device="cpu")
        tr.train_epoch(loader)
        tr.evaluate(loader)
        assert seen and all(k == "sage" for k in seen[:6]), seen
        assert loader.prepare_for is None
        n = len(seen)
        next(iter(loader))                              # outside the trainer: the caller's own setting (none)
        assert seen[n] is None
    finally:
        SubjectStore.collate = orig
