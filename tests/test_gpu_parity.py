"""The parity gate proper: the product CUDA library (lib/libcgnn.so, sm_100a) through the public API,
against golden fixtures made from the reference and against the oracle.  Run on a B200 with -m gpu."""
import numpy as np
import pytest
import torch

import helpers
import parity

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module", autouse=True)
def _native_library_loaded():
    """Fail loudly if the CUDA extension is missing - there is no fallback to hide behind."""
    assert torch.cuda.is_available()
    from connectome_gnn import _lib
    lib = _lib.load()
    before = lib.cgnn_kernel_launches()
    yield
    assert lib.cgnn_kernel_launches() > before, "no cgnn kernel was launched by the GPU tests"
    assert any("libcgnn.so" in line for line in open("/proc/self/maps")), "libcgnn.so is not mapped"


@pytest.mark.parametrize("fixture", ["ref_small.npz", "ref_ragged.npz", "ref_c1.npz", "ref_c3.npz", "ref_c5.npz"])
def test_collate_bit_exact(fixture):
    parity.check_collate(helpers.golden(fixture), DEV)


@pytest.mark.parametrize("fixture", ["ref_small.npz", "ref_ragged.npz", "ref_c1.npz", "ref_c3.npz", "ref_c5.npz"])
@pytest.mark.parametrize("kind", ["gcn", "sage"])
def test_model_parity(fixture, kind):
    """Against fixtures made from the UNMODIFIED reference (tests/golden/make_golden.py), up to 360-node subjects at
    hidden 64 and 256; every gradient tensor of the small fixtures within 1e-5 (parity.FIXTURE_TOL)."""
    parity.check_model(helpers.golden(fixture), kind, DEV, tol=parity.FIXTURE_TOL[fixture])


@pytest.mark.parametrize("kind", ["gcn", "sage"])
def test_trainer_trajectory(kind):
    parity.check_trainer(kind, DEV)


@pytest.mark.parametrize("kind", ["gcn", "sage"])
@pytest.mark.parametrize("shape", [(6, 360, 64, 3), (5, 70, 20, 2), (3, 100, 80, 2), (40, 33, 32, 3)])
def test_oracle_random_weights(kind, shape):
    """Fresh weights, package vs oracle/port.py: the 360-node / H=64 shape of BASELINE configs[2], a chunk
    boundary inside a subject, a wide layer, many small subjects."""
    from connectome_gnn.synthetic import generate_dataset
    subjects, regions, hidden, layers = shape
    graphs = generate_dataset(num_subjects=subjects, num_regions=regions, seed=11)
    parity.check_against_oracle(graphs, kind, DEV, hidden=hidden, layers=layers)


@pytest.mark.parametrize("shape", [(4, 360, 256, 3), (7, 84, 256, 2), (3, 45, 256, 3)])
def test_wide_gcn_hidden_256(shape):
    """BASELINE configs[4] shape class (GCN, hidden 256): first-layer kernel at 256 outputs, then the wide layers
    (slab gather + K-looped tcgen05 contractions, wide_tc.cu), forward, backward and BatchNorm statistics."""
    from connectome_gnn.synthetic import generate_dataset
    subjects, regions, hidden, layers = shape
    graphs = generate_dataset(num_subjects=subjects, num_regions=regions, seed=5)
    parity.check_against_oracle(graphs, "gcn", DEV, hidden=hidden, layers=layers)


def test_wide_gcn_dropout_masks_are_consistent():
    """Hidden 256 with dropout: the same seed gives the same loss twice, and the step is finite."""
    from connectome_gnn.graph import collate_graphs
    from connectome_gnn.models import GCNConnectome
    from connectome_gnn.synthetic import generate_dataset
    from connectome_gnn.train import CrossEntropyLoss
    b = collate_graphs(generate_dataset(num_subjects=5, num_regions=84, seed=2))
    torch.manual_seed(0)
    m = GCNConnectome(in_channels=5, hidden_dim=256, num_classes=2, num_layers=3, dropout=0.3).cuda().train()
    losses = []
    for _ in range(2):
        m.zero_grad()
        torch.manual_seed(123)
        loss = CrossEntropyLoss()(m(b), b.labels)
        loss.backward()
        losses.append(float(loss))
        assert all(torch.isfinite(p.grad).all() for p in m.parameters())
    assert losses[0] == losses[1]


@pytest.mark.parametrize("kind,hidden,regions", [("gcn", 64, 77), ("sage", 64, 77), ("gcn", 256, 45), ("sage", 256, 45), ("gcn", 64, 360), ("sage", 64, 360)])
def test_no_writes_outside_the_output_tensors(kind, hidden, regions, monkeypatch):
    """Every tensor the engine allocates for a call is placed between two guard bands; after a full training step and
    an eval forward over ragged row counts (not multiples of the 128-row / 16-row tiles) the bands must be untouched."""
    from connectome_gnn import _engine
    from connectome_gnn.graph import collate_graphs
    from connectome_gnn.models import GCNConnectome, GraphSAGEConnectome
    from connectome_gnn.synthetic import generate_dataset
    from connectome_gnn.train import CrossEntropyLoss
    PAD = 256                                     # elements on either side (keeps 16-byte alignment for every dtype)
    guards = []

    def guarded_empty(self, shape, dtype=torch.float32):
        shape = tuple(shape) if isinstance(shape, (tuple, list, torch.Size)) else (int(shape),)
        n = 1
        for d in shape:
            n *= int(d)
        raw = torch.empty(n + 2 * PAD, dtype=dtype, device=self.device)
        fill = 0x5A if not dtype.is_floating_point else None
        if fill is None:
            raw.fill_(float("nan"))
        else:
            raw.fill_(fill)
        guards.append((raw, n, dtype))
        return raw[PAD:PAD + n].view(shape)

    monkeypatch.setattr(_engine.Engine, "empty", guarded_empty)
    graphs = generate_dataset(num_subjects=5, num_regions=regions, seed=3) + generate_dataset(num_subjects=2, num_regions=max(8, regions // 3), seed=4)
    b = collate_graphs(graphs)
    torch.manual_seed(0)
    cls = GCNConnectome if kind == "gcn" else GraphSAGEConnectome
    m = cls(in_channels=5, hidden_dim=hidden, num_classes=2, num_layers=3, dropout=0.1).cuda().train()
    CrossEntropyLoss()(m(b), b.labels).backward()
    m.eval()
    with torch.no_grad():
        m(b)
    torch.cuda.synchronize()
    assert len(guards) > 20
    for raw, n, dtype in guards:
        for band in (raw[:PAD], raw[PAD + n:]):
            ok = torch.isnan(band).all() if dtype.is_floating_point else (band == 0x5A).all()
            assert bool(ok), f"guard band of a {dtype} tensor with {n} elements was written"


@pytest.mark.parametrize("shape", [(4, 360, 256, 3), (7, 84, 256, 2)])
def test_wide_sage_hidden_256(shape):
    """GraphSAGE at hidden 256: gather (+ transformed rows) -> K = 512 contraction in place; backward through the dz
    pass, the two weight-gradient halves, the two input-gradient halves and the transposed gather."""
    from connectome_gnn.synthetic import generate_dataset
    subjects, regions, hidden, layers = shape
    graphs = generate_dataset(num_subjects=subjects, num_regions=regions, seed=6)
    parity.check_against_oracle(graphs, "sage", DEV, hidden=hidden, layers=layers)


def test_hidden_192_fails_loudly_for_large_subjects():
    """Widths the kernels do not cover (here 192 with 360-node subjects) are an error, never a silent fallback."""
    from connectome_gnn import _lib
    from connectome_gnn.graph import collate_graphs
    from connectome_gnn.models import GraphSAGEConnectome
    from connectome_gnn.synthetic import generate_dataset
    b = collate_graphs(generate_dataset(num_subjects=2, num_regions=360, seed=1))
    m = GraphSAGEConnectome(in_channels=5, hidden_dim=192).cuda().eval()
    with pytest.raises(_lib.CgnnError), torch.no_grad():
        m(b)


def test_csr_from_coo_equals_collate():
    from connectome_gnn.graph import ConnectomeBatch, collate_graphs
    a = helpers.golden("ref_ragged.npz")
    b = collate_graphs(helpers.graphs_from_store(a))
    moved = ConnectomeBatch(b.node_features.cpu(), b.edge_index.cpu(), b.edge_weight.cpu(), b.batch.cpu(),
                            b.labels.cpu(), b.ptr.cpu()).to("cuda")
    c = moved.ensure_csr()
    for f in ("in_rowptr", "in_col", "in_w", "in_wn", "out_rowptr", "out_col", "out_w", "out_wn", "deg", "dinv", "wsum", "eptr"):
        assert torch.equal(getattr(c, f), getattr(b.csr, f)), f


def test_reference_style_usage_runs_unchanged():
    """The calls the reference's tests / README make (CPU-constructed model, device='cpu' Trainer) work as is."""
    from connectome_gnn import (ConnectomeDataLoader, GCNConnectome, GraphSAGEConnectome, Trainer, collate_graphs,
                                generate_dataset)
    graphs = generate_dataset(num_subjects=8, num_regions=20, seed=0)
    batch = collate_graphs(graphs)
    assert batch.node_features.shape == (160, 5) and batch.num_graphs == 8 and int(batch.ptr[-1]) == 160
    for cls in (GCNConnectome, GraphSAGEConnectome):
        model = cls(in_channels=5, hidden_dim=32, num_classes=2)
        model.eval()
        with torch.no_grad():
            assert model(batch).shape == (8, 2) and model.encode(batch).shape == (8, 32)
        model.train()
        model(batch).sum().backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
    graphs = generate_dataset(num_subjects=40, num_regions=20, seed=7)
    tl = ConnectomeDataLoader(graphs[:30], batch_size=10, shuffle=True)
    vl = ConnectomeDataLoader(graphs[30:], batch_size=10, shuffle=False)
    assert len(tl) == 3 and sum(b.num_graphs for b in vl) == 10
    model = GCNConnectome(in_channels=5, hidden_dim=16, num_classes=2)
    trainer = Trainer(model, torch.optim.Adam(model.parameters(), lr=1e-3), device="cpu")
    hist = trainer.fit(tl, vl, num_epochs=3, patience=10, verbose=False)
    assert len(hist["train_loss"]) == 3 and trainer.evaluate(vl)["total"] == 10


def test_dropout_statistics_and_determinism():
    from connectome_gnn import _engine
    from connectome_gnn._engine import Act
    eng = _engine.engine_for(torch.zeros(1, device=DEV))
    n, C = 360 * 64, 64
    ones = torch.ones(n, C, device=DEV)
    ptr = torch.tensor([0, n], dtype=torch.int64, device=DEV)
    one, zero = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
    for p in (0.1, 0.3, 0.5):
        emb = eng.pool_fwd(ones, Act(one, zero, False, p, seed=7, site=2), ptr, 1)
        emb2 = eng.pool_fwd(ones, Act(one, zero, False, p, seed=7, site=2), ptr, 1)
        other = eng.pool_fwd(ones, Act(one, zero, False, p, seed=8, site=2), ptr, 1)
        assert torch.equal(emb, emb2) and not torch.equal(emb, other)
        per_channel_keep = emb[0] * (1 - p)
        assert float((per_channel_keep - (1 - p)).abs().max()) < 0.02          # every channel close to 1 - p
        assert abs(float(per_channel_keep.mean()) - (1 - p)) < 2e-3


def test_dropout_gradient_consistency():
    """Dropout on: the same seed gives the same masks twice, and the analytic gradient (backward regenerates the forward
    masks) agrees with a central finite difference taken with the masks held fixed.  The probe runs along the gradient
    itself (directional derivative = |g|, no cancellation between parameters) for three mask instances."""
    from connectome_gnn.graph import collate_graphs
    from connectome_gnn.train import CrossEntropyLoss
    a = helpers.golden("ref_small.npz")
    b = collate_graphs(helpers.graphs_from_store(a))
    for kind in ("gcn", "sage"):
        m = parity.make_model(kind, a, DEV, dropout=0.3)
        m.train()
        for bn in m.batch_norms:
            bn.momentum = 0.0
        params = list(m.parameters())
        good = 0
        for seed in (123, 124, 125):
            def loss_at():
                torch.manual_seed(seed)
                return CrossEntropyLoss()(m(b), b.labels)

            m.zero_grad()
            l0 = loss_at()
            l0.backward()
            assert float(loss_at().detach()) == float(l0.detach())
            norm = float(torch.sqrt(sum((p.grad.double() ** 2).sum() for p in params)))
            direction = [(p.grad / norm).clone() for p in params]
            eps = 2e-4
            with torch.no_grad():
                for p, d in zip(params, direction): p.add_(eps * d)
                lp = float(loss_at())
                for p, d in zip(params, direction): p.sub_(2 * eps * d)
                lm = float(loss_at())
                for p, d in zip(params, direction): p.add_(eps * d)
            fd = (lp - lm) / (2 * eps)
            # GraphSAGE (ReLU inside the layer, then BatchNorm, then dropout) has kinks close enough to this step that the
            # difference quotient sits up to ~8 % under |g| for some masks and closes in as the step shrinks - identically
            # for the tensor-core and the generic kernels (tools/debug/fd_probe.py); GCN agrees to 1e-4.
            assert fd == pytest.approx(norm, rel=0.12 if kind == "sage" else 0.01), (kind, seed, fd, norm)
            good += abs(fd - norm) <= 0.05 * norm
        assert good >= 2, (kind, good)


def test_unsupported_shapes_fail_loudly():
    """No silent fallback: a tile that cannot fit shared memory raises CgnnError."""
    from connectome_gnn import _lib
    from connectome_gnn.graph import collate_graphs
    from connectome_gnn.models import GCNConnectome
    from connectome_gnn.synthetic import generate_dataset
    b = collate_graphs(generate_dataset(num_subjects=2, num_regions=360, seed=1))
    m = GCNConnectome(in_channels=5, hidden_dim=512).cuda().eval()
    with pytest.raises(_lib.CgnnError), torch.no_grad():
        m(b)


@pytest.mark.parametrize("rows,K,N", [(128, 64, 64), (360, 64, 64), (1000, 64, 64), (77, 32, 32), (300, 128, 64), (256, 64, 128)])
def test_tensor_core_projection_is_fp32_grade(rows, K, N):
    """cgnn_project_tf32x3 (tcgen05 3xTF32, accumulators in TMEM) against an fp64 matmul: error at the level of
    an fp32 FMA chain, far inside the 1e-5 budget (plain TF32 would sit at ~5e-4).  The A operand goes to tensor memory
    (tcgen05.st + the [a_tmem] form of tcgen05.mma) whenever its columns fit next to the accumulators."""
    from connectome_gnn import _engine
    eng = _engine.engine_for(torch.zeros(1, device=DEV))
    g = torch.Generator().manual_seed(rows + K + N)
    X = torch.randn(rows, K, generator=g)
    W = torch.randn(N, K, generator=g) * 0.2
    P = eng.project_tf32x3(X.to(DEV), W.to(DEV)).cpu()
    ref = (X.double() @ W.double().T)
    err = helpers.max_rel(P, ref)
    fp32 = helpers.max_rel(X @ W.T, ref)
    assert err <= 2e-6, (err, fp32)


@pytest.mark.parametrize("shape", [(6, 360, 5, 64), (6, 360, 64, 64), (9, 84, 64, 64), (5, 100, 32, 32), (4, 77, 64, 128), (3, 50, 128, 64)])
@pytest.mark.parametrize("drop", [0.0, 0.3])
def test_tensor_core_layer_kernels_match_the_generic_kernels(shape, drop):
    """Every tcgen05/TMEM layer kernel against the generic SIMT kernel of the same entry point on the same device
    inputs (CGNN_OPT_TENSOR_CORES toggled): identical up to fp32 summation order, with and without dropout."""
    from connectome_gnn import _engine, _lib
    from connectome_gnn._engine import Act
    from connectome_gnn.graph import collate_graphs
    from connectome_gnn.synthetic import generate_dataset
    subjects, regions, d_in, H = shape
    eng = _engine.engine_for(torch.zeros(1, device=DEV))
    b = collate_graphs(generate_dataset(num_subjects=subjects, num_regions=regions, seed=5))
    g = torch.Generator().manual_seed(1)
    rows = b.num_nodes
    t_in = torch.randn(rows, d_in, generator=g).to(DEV)
    W = (torch.randn(H, d_in, generator=g) * 0.2).to(DEV)
    bias = (torch.randn(H, generator=g) * 0.1).to(DEV)
    act = Act((1 + 0.1 * torch.randn(d_in, generator=g)).to(DEV), (0.1 * torch.randn(d_in, generator=g)).to(DEV),
              True, drop, seed=77, site=1, row_base=1000)
    out = {}
    for use_tc in (1, 0):
        assert eng.lib.cgnn_set_option(1, use_tc) == 0
        try:
            out[use_tc] = eng.layer_fwd("gcn", t_in, act, W, bias, b.csr, b.ptr, b.num_graphs, True)
        finally:
            eng.lib.cgnn_set_option(1, 1)
    helpers.assert_close(out[1][0], out[0][0], "gcn fwd z: tensor-core vs generic", tol=2e-6)
    assert float(out[1][1][0]) == float(out[0][1][0]) == rows
    helpers.assert_close(out[1][1][1:1 + H], out[0][1][1:1 + H], "BN mean", tol=2e-6)
    helpers.assert_close(out[1][1][1 + H:], out[0][1][1 + H:], "BN M2", tol=2e-5)


@pytest.mark.parametrize("shape", [(6, 360, 5, 64), (6, 360, 64, 64), (9, 84, 64, 64), (5, 100, 32, 32), (4, 77, 64, 128), (3, 50, 32, 64)])
@pytest.mark.parametrize("drop", [0.0, 0.3])
def test_tensor_core_sage_forward_matches_the_generic_kernel(shape, drop):
    """GraphSAGE forward as gather kernel + tcgen05 contraction against the generic SIMT kernel (same entry point)."""
    from connectome_gnn import _engine
    from connectome_gnn._engine import Act
    from connectome_gnn.graph import collate_graphs
    from connectome_gnn.synthetic import generate_dataset
    subjects, regions, d_in, H = shape
    eng = _engine.engine_for(torch.zeros(1, device=DEV))
    b = collate_graphs(generate_dataset(num_subjects=subjects, num_regions=regions, seed=6))
    g = torch.Generator().manual_seed(2)
    rows = b.num_nodes
    t_in = torch.randn(rows, d_in, generator=g).to(DEV)
    W = (torch.randn(H, 2 * d_in, generator=g) * 0.2).to(DEV)
    bias = (torch.randn(H, generator=g) * 0.1).to(DEV)
    act = Act((1 + 0.1 * torch.randn(d_in, generator=g)).to(DEV), (0.1 * torch.randn(d_in, generator=g)).to(DEV),
              False, drop, seed=78, site=2, row_base=500)
    out = {}
    for use_tc in (1, 0):
        assert eng.lib.cgnn_set_option(1, use_tc) == 0
        try:
            out[use_tc] = eng.layer_fwd("sage", t_in, act, W, bias, b.csr, b.ptr, b.num_graphs, True)
        finally:
            eng.lib.cgnn_set_option(1, 1)
    helpers.assert_close(out[1][0], out[0][0], "sage fwd z: tensor-core vs generic", tol=2e-6)
    assert float(out[1][1][0]) == float(out[0][1][0]) == rows
    helpers.assert_close(out[1][1][1:1 + H], out[0][1][1:1 + H], "BN mean", tol=2e-6)
    helpers.assert_close(out[1][1][1 + H:], out[0][1][1 + H:], "BN M2", tol=2e-5)


@pytest.mark.parametrize("kind", ["gcn", "sage"])
@pytest.mark.parametrize("shape", [(6, 360, 5, 64), (6, 360, 64, 64), (9, 84, 64, 64), (5, 100, 32, 32), (4, 77, 64, 128), (3, 50, 32, 64)])
@pytest.mark.parametrize("top", [False, True])
def test_tensor_core_backward_matches_the_generic_kernels(kind, shape, top):
    """Layer backward (gather kernel + tcgen05 contractions) against the generic SIMT kernels: dW, dbias, du_in and
    the BatchNorm-backward sums of the layer below, for a middle layer (per-row upstream) and the top layer
    (pooled upstream), with dropout on both sides."""
    from connectome_gnn import _engine
    from connectome_gnn._engine import Act, BnBwd
    from connectome_gnn.graph import collate_graphs
    from connectome_gnn.synthetic import generate_dataset
    subjects, regions, d_in, H = shape
    eng = _engine.engine_for(torch.zeros(1, device=DEV))
    b = collate_graphs(generate_dataset(num_subjects=subjects, num_regions=regions, seed=7))
    g = torch.Generator().manual_seed(3)
    rows, B = b.num_nodes, b.num_graphs
    rn = lambda *s: torch.randn(*s, generator=g)
    t_in = rn(rows, d_in).to(DEV)
    W = (rn(H, d_in if kind == "gcn" else 2 * d_in) * 0.2).to(DEV)
    bias = (rn(H) * 0.1).to(DEV)
    act_in = Act((1 + 0.1 * rn(d_in)).to(DEV), (0.1 * rn(d_in)).to(DEV), kind == "gcn", 0.3, seed=79, site=0, row_base=64)
    assert eng.lib.cgnn_set_option(1, 0) == 0
    try:
        z, _, agg = eng.layer_fwd(kind, t_in, act_in, W, bias, b.csr, b.ptr, B, False)
    finally:
        eng.lib.cgnn_set_option(1, 1)
    act_out = Act((1 + 0.1 * rn(H)).to(DEV), (0.1 * rn(H)).to(DEV), kind == "gcn", 0.3, seed=79, site=1, row_base=64)
    mean, rstd = (0.1 * rn(H)).to(DEV), (1 + 0.1 * rn(H)).abs().to(DEV)
    sums = (rn(2, H) * 0.5).to(DEV)
    bn = BnBwd(act_out.scale, mean, rstd, sums, float(rows), True)
    du = None if top else rn(rows, H).to(DEV)
    demb = rn(B, H).to(DEV) if top else None
    pmean, prstd = (0.1 * rn(d_in)).to(DEV), (1 + 0.1 * rn(d_in)).abs().to(DEV)
    out = {}
    for use_tc in (1, 0):
        assert eng.lib.cgnn_set_option(1, use_tc) == 0
        try:
            out[use_tc] = eng.layer_bwd(kind, du, demb, z, act_out, bn, t_in, act_in, W, b.csr, b.ptr, B, True, pmean, prstd, agg)
        finally:
            eng.lib.cgnn_set_option(1, 1)
    for name, got, ref in zip(("dW", "dbias", "du_in", "prev_sums"), out[1], out[0]):
        helpers.assert_close(got, ref, f"{kind} bwd {name}: tensor-core vs generic", tol=5e-6)
    # the float64 record of the sums: written next to the fp32 one, and read instead of it when handed in
    for use_tc in (1, 0):
        assert torch.equal(out[use_tc][4].float(), out[use_tc][3]), "prev_sums64 must round to prev_sums"
    bn64 = BnBwd(act_out.scale, mean, rstd, torch.full_like(sums, float("nan")), float(rows), True, sums.double())
    got64 = eng.layer_bwd(kind, du, demb, z, act_out, bn64, t_in, act_in, W, b.csr, b.ptr, B, True, pmean, prstd, agg)
    for name, got, ref in zip(("dW", "dbias", "du_in", "prev_sums"), got64, out[1]):
        helpers.assert_close(got, ref, f"{kind} bwd {name}: sums read from the float64 record", tol=1e-6)


def test_streaming_store_matches_resident_store():
    """Double-buffered host->device staging (uploads on a side stream) collates exactly what a resident store does."""
    from connectome_gnn.graph import StreamingStore, SubjectStore, pack_graphs
    from connectome_gnn.synthetic import generate_dataset
    sets = [pack_graphs(generate_dataset(num_subjects=12, num_regions=40, seed=s)) for s in (11, 12, 13)]
    pinned = [{k: (v.pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in p.items()} for p in sets]
    ss = StreamingStore(pinned[0], DEV)
    ids = np.arange(12)[::-1].copy()
    ss.prefetch(pinned[0])
    for k in range(3):
        st = ss.next()
        if k + 1 < 3:
            ss.prefetch(pinned[k + 1])
        got = st.collate(ids)
        ref = SubjectStore(sets[k], DEV).collate(ids)
        for f in ("node_features", "edge_index", "edge_weight", "batch", "labels", "ptr"):
            assert torch.equal(getattr(got, f), getattr(ref, f)), f
        assert torch.equal(got.csr.deg, ref.csr.deg) and torch.equal(got.csr.in_col, ref.csr.in_col)
        torch.cuda.synchronize()
    with pytest.raises(RuntimeError):
        ss.next()


def test_compact_store_collates_identically():
    """pack_graphs(compact=True) (src | dst << 16 in one int32, dst pointer NULL) gives the same batch bit for bit."""
    from connectome_gnn.graph import SubjectStore, pack_graphs
    from connectome_gnn.synthetic import generate_dataset
    graphs = generate_dataset(num_subjects=9, num_regions=84, seed=21)
    ids = np.array([3, 0, 8, 5, 5, 1])
    a = SubjectStore(pack_graphs(graphs), DEV).collate(ids)
    b = SubjectStore(pack_graphs(graphs, compact=True), DEV).collate(ids)
    for f in ("node_features", "edge_index", "edge_weight", "batch", "labels", "ptr"):
        assert torch.equal(getattr(a, f), getattr(b, f)), f
    for f in ("in_rowptr", "in_col", "in_w", "in_wn", "out_rowptr", "out_col", "out_w", "out_wn", "deg", "dinv", "wsum"):
        assert torch.equal(getattr(a.csr, f), getattr(b.csr, f)), f


def test_pair_store_collates_identically():
    """pack_graphs(compact=True, pairs=True): one entry per undirected edge (edge_pairs = 1) expands to the same batch,
    CSR and aggregation blobs bit for bit; a list that is not made of reversed pairs keeps the plain compact form."""
    from connectome_gnn.graph import ConnectomeGraph, SubjectStore, pack_graphs
    from connectome_gnn.synthetic import generate_dataset
    graphs = generate_dataset(num_subjects=6, num_regions=84, seed=21) + generate_dataset(num_subjects=3, num_regions=360, seed=22)
    ids = np.array([3, 0, 8, 5, 5, 1, 7])
    packed = pack_graphs(graphs, compact=True, pairs=True)
    assert packed["edge_pairs"] == 1 and packed["src"].numel() * 2 == int(packed["edge_ptr"][-1])
    for kind in ("gcn", "sage"):
        a = SubjectStore(pack_graphs(graphs), DEV).collate(ids, prepare_for=kind)
        b = SubjectStore(packed, DEV).collate(ids, prepare_for=kind)
        for f in ("node_features", "edge_index", "edge_weight", "batch", "labels", "ptr"):
            assert torch.equal(getattr(a, f), getattr(b, f)), f
        for f in ("in_rowptr", "in_col", "in_w", "in_wn", "out_rowptr", "out_col", "out_w", "out_wn", "deg", "dinv", "wsum"):
            assert torch.equal(getattr(a.csr, f), getattr(b.csr, f)), f
        # the aggregation blobs (uninitialised slack between subjects, so not comparable as buffers): same logits
        from connectome_gnn.models import GCNConnectome, GraphSAGEConnectome
        torch.manual_seed(0)
        m = (GCNConnectome if kind == "gcn" else GraphSAGEConnectome)(in_channels=5, hidden_dim=64).cuda().eval()
        with torch.no_grad():
            assert torch.equal(m(a), m(b))
    g0 = graphs[0]
    odd = ConnectomeGraph(g0.node_features, g0.edge_index[:, :-1], g0.edge_weight[:-1], g0.label)   # one edge without its reverse
    plain = pack_graphs([odd] + graphs[1:], compact=True, pairs=True)
    assert plain["edge_pairs"] == 0 and plain["src"].numel() == int(plain["edge_ptr"][-1])
    c = SubjectStore(plain, DEV).collate(np.array([0, 2]))
    d = SubjectStore(pack_graphs([odd] + graphs[1:]), DEV).collate(np.array([0, 2]))
    assert torch.equal(c.edge_index, d.edge_index) and torch.equal(c.csr.in_wn, d.csr.in_wn)


@pytest.mark.parametrize("kind", ["gcn", "sage"])
def test_no_reads_of_unwritten_memory(kind):
    """The allocator's free blocks are filled with zeros, NaN and huge values between runs (torch.empty then hands the
    kernels exactly that garbage): inference on both paths, lean and full batches, and a training step give the same
    bits every time - nothing reads a buffer before it is written (compute-sanitizer's initcheck is closed on this pool)."""
    from connectome_gnn.graph import SubjectStore, pack_graphs
    from connectome_gnn.models import GCNConnectome, GraphSAGEConnectome
    from connectome_gnn.synthetic import generate_dataset
    from connectome_gnn.train import CrossEntropyLoss
    store = SubjectStore(pack_graphs(generate_dataset(num_subjects=12, num_regions=360, seed=42)), DEV)

    def poison(value):
        blocks = [torch.full((n,), value, device=DEV) for n in (1 << 25, 1 << 23, 1 << 21, 1 << 19, 1 << 17, 1 << 15, 1 << 13) for _ in range(3)]
        torch.cuda.synchronize()
        del blocks

    torch.manual_seed(0)
    cls = GCNConnectome if kind == "gcn" else GraphSAGEConnectome
    m = cls(in_channels=5, hidden_dim=64, num_classes=2, num_layers=3, dropout=0.0).to(DEV)
    for fused in (False, True):
        m.eval()
        m.fused_eval = fused
        outs = []
        for value in (0.0, float("nan"), 3e38):
            poison(value)
            with torch.no_grad():
                outs.append(torch.cat([m(store.collate(np.arange(12), prepare_for=kind)), m(store.collate(np.array([5]))),
                                       m(store.collate(np.arange(3, 9), prepare_for=kind, backward=False))]).clone())
        assert torch.isfinite(torch.stack(outs)).all()
        assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2]), (kind, fused)
    m.train()
    grads = []
    for value in (0.0, float("nan"), 3e38):
        poison(value)
        m.zero_grad()
        for bn in m.batch_norms:
            bn.reset_running_stats()
        batch = store.collate(np.arange(12), prepare_for=kind)
        CrossEntropyLoss()(m(batch), batch.labels).backward()
        grads.append(torch.cat([p.grad.reshape(-1) for p in m.parameters()]).clone())
    assert torch.isfinite(torch.stack(grads)).all()
    assert torch.equal(grads[0], grads[1]) and torch.equal(grads[0], grads[2]), kind


@pytest.mark.parametrize("kind", ["gcn", "sage"])
def test_fused_step_flat_buffers(kind):
    parity.check_fused_step(DEV, kind)


def test_adam_kernel_matches_torch_adam():
    worst, equal = parity.check_adam_kernel(DEV)
    print(f"cgnn_adam_step vs torch.optim.Adam: max-norm relative difference {worst:.2e}, bit-identical after {equal} of 6 steps")


@pytest.mark.parametrize("kind", ["gcn", "sage"])
def test_graphed_step(kind):
    """BASELINE configs[0] / [1] shape (batch 16, 84-node subjects, hidden 64) as CUDA graphs."""
    train_us, eval_us = parity.check_graphed_step(kind, batch=16, regions=84)
    print(f"{kind}: graphed train step {train_us:.0f} us, graphed eval step {eval_us:.0f} us (batch 16, 84 nodes, collate included)")


@pytest.mark.parametrize("layers", [2, 3, 4])
def test_fused_eval(layers):
    parity.check_fused_eval(DEV, layers)


@pytest.mark.parametrize("compact", [None, False])      # pair store (k_collate_pairs) / plain store (k_collate_graph)
@pytest.mark.parametrize("kind", ["gcn", "sage"])
def test_lean_collate(kind, compact):
    parity.check_lean_collate(DEV, kind, compact)


@pytest.mark.parametrize("sizes", [(30, 84, 57, 130), (360, 200, 360, 84)])
def test_pair_collate_bit_exact(sizes):
    parity.check_pair_collate(DEV, sizes)


def test_pooled_last_layer():
    parity.check_pooled_last_layer(DEV)


def test_multi_subject_units():
    parity.check_multi_subject_units(DEV)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_pair_collate_fuzz(seed):
    parity.check_pair_collate_fuzz(DEV, seed=seed, rounds=8)


@pytest.mark.parametrize("kind", ["gcn", "sage"])
def test_batches_die_by_refcount(kind):
    parity.check_batches_die_by_refcount(DEV, kind, num_regions=120)


@pytest.mark.parametrize("kind", ["gcn", "sage"])
def test_collate_emits_the_same_aggregation_structure(kind):
    """`prepare_for` (blobs written by the collate kernel) and the lazy `cgnn_build_agg` path feed the layer kernels
    identical structure: forward outputs and backward gradients are bit-identical."""
    from connectome_gnn import _engine
    from connectome_gnn._engine import Act
    from connectome_gnn.graph import SubjectStore, pack_graphs
    from connectome_gnn.synthetic import generate_dataset
    graphs = generate_dataset(num_subjects=7, num_regions=84, seed=31) + generate_dataset(num_subjects=3, num_regions=50, seed=32)
    store = SubjectStore(pack_graphs(graphs), DEV)
    ids = np.array([9, 0, 4, 7, 2, 2, 8])
    a = store.collate(ids, prepare_for=kind)
    b = store.collate(ids)
    assert a.csr.agg is not None and kind in a.csr.agg and b.csr.agg is None
    eng = _engine.engine_for(a.node_features)
    g = torch.Generator().manual_seed(5)
    t_in = torch.randn(a.num_nodes, 64, generator=g).to(DEV)
    W = (torch.randn(64, 64 if kind == "gcn" else 128, generator=g) * 0.2).to(DEV)
    bias = (torch.randn(64, generator=g) * 0.1).to(DEV)
    act = Act()
    za, _, agg_a = eng.layer_fwd(kind, t_in, act, W, bias, a.csr, a.ptr, a.num_graphs, False)
    zb, _, agg_b = eng.layer_fwd(kind, t_in, act, W, bias, b.csr, b.ptr, b.num_graphs, False)
    assert torch.equal(za, zb)
    du = torch.randn(a.num_nodes, 64, generator=g).to(DEV)
    oa = eng.layer_bwd(kind, du, None, za, Act(), None, t_in, act, W, a.csr, a.ptr, a.num_graphs, True, None, None, agg_a)
    ob = eng.layer_bwd(kind, du, None, zb, Act(), None, t_in, act, W, b.csr, b.ptr, b.num_graphs, True, None, None, agg_b)
    for x, y in zip(oa[:3], ob[:3]):
        assert torch.equal(x, y)


@pytest.mark.parametrize("kind", ["gcn", "sage"])
@pytest.mark.parametrize("shape", [(6, 360, 5, 64), (9, 84, 5, 64), (5, 100, 8, 32), (4, 77, 3, 128)])
@pytest.mark.parametrize("top", [False, True])
def test_first_layer_kernels_match_the_generic_kernels(kind, shape, top):
    """Narrow-input layer (no tensor cores, no input gradient): forward z / BatchNorm statistics and backward dW / dbias
    against the generic SIMT kernels, with an activation on the input, dropout after the layer and a ragged batch."""
    from connectome_gnn import _engine
    from connectome_gnn._engine import Act, BnBwd
    from connectome_gnn.graph import collate_graphs
    from connectome_gnn.synthetic import generate_dataset
    subjects, regions, d_in, H = shape
    eng = _engine.engine_for(torch.zeros(1, device=DEV))
    b = collate_graphs(generate_dataset(num_subjects=subjects, num_regions=regions, seed=8) +
                       generate_dataset(num_subjects=2, num_regions=max(8, regions // 2), seed=9))
    g = torch.Generator().manual_seed(4)
    rows, B = b.num_nodes, b.num_graphs
    rn = lambda *s: torch.randn(*s, generator=g)
    t_in = rn(rows, d_in).to(DEV)
    W = (rn(H, d_in if kind == "gcn" else 2 * d_in) * 0.3).to(DEV)
    bias = (rn(H) * 0.1).to(DEV)
    act_in = Act()
    fwd = {}
    for use_tc in (1, 0):
        assert eng.lib.cgnn_set_option(1, use_tc) == 0
        try:
            fwd[use_tc] = eng.layer_fwd(kind, t_in, act_in, W, bias, b.csr, b.ptr, B, True)
        finally:
            eng.lib.cgnn_set_option(1, 1)
    helpers.assert_close(fwd[1][0], fwd[0][0], f"{kind} first layer z", tol=2e-6)
    helpers.assert_close(fwd[1][1][1:1 + H], fwd[0][1][1:1 + H], "BN mean", tol=2e-6, atol=1e-6)
    helpers.assert_close(fwd[1][1][1 + H:], fwd[0][1][1 + H:], "BN M2", tol=2e-5)
    if kind == "sage":
        helpers.assert_close(fwd[1][2], fwd[0][2], "stored aggregate", tol=2e-6)
    z, agg = fwd[0][0], fwd[0][2]
    act_out = Act((1 + 0.1 * rn(H)).to(DEV), (0.1 * rn(H)).to(DEV), kind == "gcn", 0.3, seed=80, site=1, row_base=32)
    mean, rstd = (0.1 * rn(H)).to(DEV), (1 + 0.1 * rn(H)).abs().to(DEV)
    bn = BnBwd(act_out.scale, mean, rstd, (rn(2, H) * 0.5).to(DEV), float(rows), True)
    du = None if top else rn(rows, H).to(DEV)
    demb = rn(B, H).to(DEV) if top else None
    out = {}
    for use_tc in (1, 0):
        assert eng.lib.cgnn_set_option(1, use_tc) == 0
        try:
            out[use_tc] = eng.layer_bwd(kind, du, demb, z, act_out, bn, t_in, act_in, W, b.csr, b.ptr, B, False, None, None, agg)
        finally:
            eng.lib.cgnn_set_option(1, 1)
    for name, got, ref in zip(("dW", "dbias"), out[1], out[0]):
        helpers.assert_close(got, ref, f"{kind} first layer bwd {name}", tol=5e-6)


@pytest.mark.parametrize("kind,fused", [("gcn", "auto"), ("gcn", False), ("sage", False)])
def test_same_launch_same_bits(kind, fused):
    """Every forward path (warp-specialised engine, fused eval kernel, gather + contraction) and a training step's gradients
    give the same bits launch after launch - the check that caught an intermittently wrong N = 128 tcgen05.mma."""
    from connectome_gnn.graph import SubjectStore, pack_graphs
    from connectome_gnn.models import GCNConnectome, GraphSAGEConnectome
    from connectome_gnn.synthetic import generate_connectome
    from connectome_gnn.train import CrossEntropyLoss
    sizes = [360] * 6 + [84, 30, 130, 57, 200]
    store = SubjectStore(pack_graphs([generate_connectome(num_regions=n, seed=300 + k) for k, n in enumerate(sizes)]), DEV)
    ids = np.arange(len(sizes))
    torch.manual_seed(0)
    m = (GCNConnectome if kind == "gcn" else GraphSAGEConnectome)(in_channels=5, hidden_dim=64, num_classes=2, num_layers=3,
                                                                  dropout=0.25).to(DEV)
    m.fused_eval = fused
    m.eval()
    first = None
    for rep in range(24):
        with torch.no_grad():
            out = m(store.collate(ids, prepare_for=kind, backward=False))
        first = out.clone() if first is None else first
        assert torch.equal(out, first), f"{kind} eval forward changed at repeat {rep}"
    m.train()
    g0 = None
    for rep in range(8):
        m.zero_grad()
        for bn in m.batch_norms:
            bn.reset_running_stats()
        torch.manual_seed(5)
        b = store.collate(ids, prepare_for=kind)
        CrossEntropyLoss()(m(b), b.labels).backward()
        g = torch.cat([p.grad.reshape(-1) for p in m.parameters()])
        g0 = g.clone() if g0 is None else g0
        assert torch.equal(g, g0), f"{kind} training gradients changed at repeat {rep}"
