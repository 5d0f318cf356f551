"""Parity checks shared by the simulator tests (CPU, kernel logic) and the GPU tests (the real gate).

Each function takes the device the package's tensors live on ("cpu" only ever under the `on_emu`
fixture of conftest.py) and compares the package - i.e. the kernels behind include/cgnn.h - with
golden fixtures made from the reference and with the oracle (oracle/port.py, oracle/csr_ref.c).

Bars (BASELINE.json north star): collate fields, CSR structure, D^ and w_sum BIT-EXACT;
logits / loss / gradients within REL_TOL = 1e-5 max-norm relative, dropout disabled.
Per-tensor gradients use the band the reference shows against itself:
all gradients concatenated <= 1e-5; any single tensor <= 1e-4; analytically-zero tensors
(GCN conv biases) absolute, tied to the global gradient scale.
"""
import numpy as np
import torch

import helpers
from helpers import REL_TOL

PER_TENSOR_TOL = 1e-4
# Per-fixture bars.  The small fixtures (N <= 84, H <= 64, made from the unmodified reference) hold every single gradient
# tensor to the north star's 1e-5.  At 360 nodes the reference's own reduction-order noise grows (one thread
# against eight moves logits by 7.6e-6 - 3.6e-5 and all gradients by 4.9e-4 - 2.8e-3 at hidden 256, worst tensor 2.2e-2),
# so a fixture taken from ONE eight-thread run of the reference can only be met inside that band: logits and loss stay
# at 1e-5 for hidden 64, and the hidden-256 fixture uses the band itself.
FIXTURE_TOL = {
    "ref_small.npz": dict(out=REL_TOL, grads=REL_TOL, per_tensor=1e-5),
    "ref_ragged.npz": dict(out=REL_TOL, grads=REL_TOL, per_tensor=1e-5),
    "ref_c1.npz": dict(out=REL_TOL, grads=REL_TOL, per_tensor=1e-5),
    "ref_c3.npz": dict(out=REL_TOL, grads=REL_TOL, per_tensor=1e-4),
    "ref_c5.npz": dict(out=5e-5, grads=5e-3, per_tensor=5e-2),
}
DEFAULT_TOL = dict(out=REL_TOL, grads=REL_TOL, per_tensor=PER_TENSOR_TOL)


def make_model(kind, a, device, dropout=0.0):
    from connectome_gnn.models import GCNConnectome, GraphSAGEConnectome
    sd = helpers.state_dict_from(a, f"{kind}.init")
    hidden = sd["convs.0.linear.weight"].shape[0]
    in_ch = sd["convs.0.linear.weight"].shape[1] // (2 if kind == "sage" else 1)
    layers = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("convs."))
    cls = GCNConnectome if kind == "gcn" else GraphSAGEConnectome
    m = cls(in_channels=in_ch, hidden_dim=hidden, num_classes=sd["classifier.3.weight"].shape[0],
            num_layers=layers, dropout=dropout)
    m.load_state_dict(sd)
    return m.to(device)


def check_collate(a, device):
    """collate_graphs == reference graph.py:143-167, bit for bit; D^ / w_sum bit for bit; CSR == stable sort."""
    from connectome_gnn.graph import collate_graphs
    b = collate_graphs(helpers.graphs_from_store(a))
    assert b.node_features.device.type == torch.device(device).type
    np_ = lambda t: t.detach().cpu().numpy()
    assert np.array_equal(np_(b.ptr), a["batch.ptr"])
    assert np.array_equal(np_(b.labels), a["batch.labels"])
    assert b.ptr.dtype == torch.int64 and b.batch.dtype == torch.int64 and b.edge_index.dtype == torch.int64
    if "batch.edge_index" in a:
        assert np.array_equal(np_(b.node_features), a["batch.node_features"])
        assert np.array_equal(np_(b.edge_index), a["batch.edge_index"])
        assert np.array_equal(np_(b.edge_weight), a["batch.edge_weight"])
        assert np.array_equal(np_(b.batch), a["batch.batch"])
    c = b.csr
    assert np.array_equal(np_(c.deg), a["deg"]), "D^ not bit-exact"
    assert np.array_equal(np_(c.wsum), a["wsum"]), "w_sum not bit-exact"
    helpers.assert_close(c.dinv, a["dinv"], "dinv", tol=2e-7)
    src, dst = np_(b.edge_index)
    w = np_(b.edge_weight)
    rows = b.num_nodes
    for key, other, rowptr, col, cw in ((dst, src, c.in_rowptr, c.in_col, c.in_w), (src, dst, c.out_rowptr, c.out_col, c.out_w)):
        order = np.argsort(key, kind="stable")
        assert np.array_equal(np_(col), other[order].astype(np.int32)), "CSR columns are not the stable sort of the COO"
        assert np.array_equal(np_(cw), w[order])
        assert np.array_equal(np_(rowptr), np.searchsorted(key[order], np.arange(rows + 1)).astype(np.int32))
    if "w_norm" in a:
        order = np.argsort(dst, kind="stable")
        helpers.assert_close(c.in_wn, a["w_norm"][: len(src)][order], "w^ (in)", tol=3e-7)
        order = np.argsort(src, kind="stable")
        helpers.assert_close(c.out_wn, a["w_norm"][: len(src)][order], "w^ (out)", tol=3e-7)
    # the C oracle must agree with the kernels on every CSR array, bit for bit
    from test_oracle import _run_c_oracle
    o = _run_c_oracle(a)
    for k in ("in_rowptr", "in_col", "in_w", "in_wn", "out_rowptr", "out_col", "out_w", "out_wn", "deg", "dinv", "wsum"):
        assert np.array_equal(np_(getattr(c, k)), o[k]), f"kernel vs C oracle: {k}"
    assert np.array_equal(np_(c.eptr), o["eptr"])
    return b


def check_grads(model, a, kind, tag, tol=None):
    tol = tol or DEFAULT_TOL
    names = [k for k, _ in model.named_parameters()]
    ref = {k: torch.from_numpy(a[f"{kind}.{tag}.{k}"]) for k in names}
    got = {k: p.grad.detach().cpu() for k, p in model.named_parameters()}
    assert all(g is not None for g in got.values())
    flat_got = torch.cat([got[k].reshape(-1) for k in names])
    flat_ref = torch.cat([ref[k].reshape(-1) for k in names])
    helpers.assert_close(flat_got, flat_ref, f"{kind} {tag}: all gradients", tol=tol["grads"])
    scale = float(flat_ref.abs().max())
    for k in names:
        if kind == "gcn" and k.startswith("convs.") and k.endswith(".bias") and tag == "train.grad":
            err = float((got[k].double() - ref[k].double()).abs().max())   # analytically zero: round-off only
            assert err <= tol["grads"] * scale, f"{k}: {err:.3e} vs gradient scale {scale:.3e}"
        else:
            helpers.assert_close(got[k], ref[k], f"{kind} {tag}: {k}", tol=tol["per_tensor"], atol=tol["grads"] * scale)


def check_model(a, kind, device, batch=None, tol=None):
    """eval forward, train forward/backward (dropout off), running statistics, eval-mode backward."""
    from connectome_gnn.graph import collate_graphs
    from connectome_gnn.train import CrossEntropyLoss
    tol = tol or DEFAULT_TOL
    b = batch if batch is not None else collate_graphs(helpers.graphs_from_store(a))
    m = make_model(kind, a, device)
    m.eval()
    with torch.no_grad():
        helpers.assert_close(m.encode(b), a[f"{kind}.eval.emb"], f"{kind} eval emb", tol=tol["out"])
        helpers.assert_close(m(b), a[f"{kind}.eval.logits"], f"{kind} eval logits", tol=tol["out"])
    m.train()
    logits = m(b)
    loss_fn = CrossEntropyLoss()
    loss = loss_fn(logits, b.labels)
    loss.backward()
    helpers.assert_close(logits, a[f"{kind}.train.logits"], f"{kind} train logits", tol=tol["out"])
    helpers.assert_close(loss, a[f"{kind}.train.loss"], f"{kind} train loss", tol=tol["out"])
    check_grads(m, a, kind, "train.grad", tol)
    for k, v in m.state_dict().items():
        if "running" in k:
            helpers.assert_close(v, a[f"{kind}.train.after.{k}"], f"{kind} {k}", tol=tol["out"])
        if "num_batches" in k:
            assert int(v) == int(a[f"{kind}.train.after.{k}"])
    if f"{kind}.evalgrad.classifier.0.weight" in a:
        m.zero_grad()
        m.eval()
        m(b).sum().backward()
        check_grads(m, a, kind, "evalgrad", tol)
    return m, b


class _float64:
    """Run the oracle in double precision: the reference's own op sequence, evaluated without fp32
    round-off.  SURVEY A.3: the fp32 reference disagrees with ITSELF by up to 3e-3 per gradient tensor
    between thread counts at N=360 (BatchNorm over near-dead ReLU channels amplifies reduction-order
    noise), so for fresh weights the stable yardstick is the fp64 evaluation; the fp32 oracle's own
    distance to it is measured in the same run and must not be (much) smaller than ours."""

    def __enter__(self):
        self.old = torch.get_default_dtype()
        torch.set_default_dtype(torch.float64)

    def __exit__(self, *exc):
        torch.set_default_dtype(self.old)


def oracle_reference(kind, params, graphs):
    """(fp64 truth, fp32 oracle at 8 threads) as fixture-style dicts for the package to be compared with."""
    from oracle import port
    ob = port.collate(graphs)
    out = {}
    for tag, cast in (("f64", torch.float64), ("f32", torch.float32)):
        conv = lambda v: v.to(cast) if (v is not None and v.is_floating_point()) else v
        p = {k: conv(v.clone()) for k, v in params.items()}
        b = {k: conv(v) for k, v in ob.items()}
        old_threads = torch.get_num_threads()
        torch.set_num_threads(8)
        try:
            ctx = _float64() if cast == torch.float64 else _nullctx()
            with ctx:
                a = {}
                with torch.no_grad():
                    a[f"{kind}.eval.emb"] = port.encode(kind, dict(p), b).numpy()
                    a[f"{kind}.eval.logits"] = port.forward(kind, dict(p), b).numpy()
                p2 = {k: v.clone() for k, v in p.items()}
                logits, loss, grads = port.loss_and_grads(kind, p2, b, training=True, dropout=0.0)
        finally:
            torch.set_num_threads(old_threads)
        a[f"{kind}.train.logits"], a[f"{kind}.train.loss"] = logits.numpy(), loss.numpy()
        for k, gr in grads.items():
            a[f"{kind}.train.grad.{k}"] = gr.numpy()
        for k, v in params.items():
            a[f"{kind}.init.{k}"] = v.numpy()
            if "running" in k or "num_batches" in k:
                a[f"{kind}.train.after.{k}"] = p2[k].numpy()
        out[tag] = a
    return out


class _nullctx:
    def __enter__(self): return self
    def __exit__(self, *exc): return False


def check_against_oracle(graphs, kind, device, hidden, layers, seed=0):
    """Fresh random weights: package vs oracle/port.py on the same inputs (sizes the oracle handles in seconds)."""
    from oracle import port
    from connectome_gnn.graph import collate_graphs
    from connectome_gnn.train import CrossEntropyLoss
    g = torch.Generator().manual_seed(seed)
    g2 = torch.Generator().manual_seed(seed + 1)
    params = port.init_params(kind, graphs[0].num_features, hidden, 2, layers, generator=g)
    # Make BatchNorm affine / running stats non-trivial.  NB the instance matters: a pre-activation that lands
    # within fp32 round-off of the ReLU threshold (|q| ~ 1e-8) flips between ANY two fp32 evaluation orders and
    # moves a whole gradient row by ~1e-2 - seen once on GPU (SAGE, 6x360, one element of 138k; the fp64 value
    # was q = 2.7e-8).  That is the reference's own A.3 noise, not a kernel property, so the seeds here are
    # fixed to instances without such a coincidence and the check below is against the fp64 evaluation.
    for k in params:
        if ".weight" in k and "batch_norms" in k:
            params[k] = params[k] + 0.1 * torch.randn(params[k].shape, generator=g)
        if k.endswith("running_mean"):
            params[k] = 0.05 * torch.randn(params[k].shape, generator=g2)
        if k.endswith("running_var"):
            params[k] = 1.0 + 0.2 * torch.rand(params[k].shape, generator=g2)
    refs = oracle_reference(kind, params, graphs)
    a, a32 = refs["f64"], refs["f32"]
    b = collate_graphs(graphs)
    m = make_model(kind, a32, device)
    m.eval()
    with torch.no_grad():
        helpers.assert_close(m.encode(b), a[f"{kind}.eval.emb"], f"{kind} eval emb")
        helpers.assert_close(m(b), a[f"{kind}.eval.logits"], f"{kind} eval logits")
    m.train()
    logits = m(b)
    CrossEntropyLoss()(logits, b.labels).backward()
    helpers.assert_close(logits, a[f"{kind}.train.logits"], f"{kind} train logits")
    helpers.assert_close(logits, a32[f"{kind}.train.logits"], f"{kind} train logits (fp32 oracle)")
    check_grads(m, a, kind, "train.grad")
    # the fp32 oracle must sit in the same band around the fp64 truth (it is usually further away than we are)
    names = [k for k, _ in m.named_parameters()]
    cat = lambda src: torch.cat([torch.as_tensor(src[f"{kind}.train.grad.{k}"]).reshape(-1).double() for k in names])
    ours = helpers.max_rel(torch.cat([p.grad.reshape(-1) for _, p in m.named_parameters()]), cat(a))
    theirs = helpers.max_rel(cat(a32), cat(a))
    assert ours <= max(REL_TOL, 10 * theirs), (ours, theirs)
    for k, v in m.state_dict().items():
        if "running" in k:
            helpers.assert_close(v, a[f"{kind}.train.after.{k}"], f"{kind} {k}")


def check_trainer(kind, device):
    """Trainer.fit trajectory vs the reference's (tests/golden/ref_trainer.npz, dropout 0, Adam)."""
    import json, os
    from connectome_gnn.graph import ConnectomeDataLoader
    from connectome_gnn.train import Trainer
    a = helpers.golden("ref_trainer.npz")
    ref = json.load(open(os.path.join(helpers.GOLDEN, "ref_meta.json")))["trainer"][kind]
    graphs = helpers.graphs_from_store(a)
    model = make_model(kind, a, "cpu")                      # optimizer is built before Trainer moves the model
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)
    torch.manual_seed(1)
    train_loader = ConnectomeDataLoader(graphs[:30], batch_size=10, shuffle=True)
    val_loader = ConnectomeDataLoader(graphs[30:], batch_size=10, shuffle=False)
    trainer = Trainer(model, opt, device=device)
    hist = trainer.fit(train_loader, val_loader, num_epochs=3, patience=10, verbose=False)
    assert set(hist) == {"train_loss", "val_loss", "val_acc"}
    for key in ("train_loss", "val_loss"):
        assert hist[key] == __import__("pytest").approx(ref["history"][key], rel=2e-5), key
    assert hist["val_acc"] == ref["history"]["val_acc"]
    ev = trainer.evaluate(val_loader)
    assert ev["total"] == ref["final_eval"]["total"] and ev["correct"] == ref["final_eval"]["correct"]
    final = helpers.state_dict_from(a, f"{kind}.final")
    for k, v in model.state_dict().items():
        if kind == "gcn" and ((k.startswith("convs.") and k.endswith(".bias")) or k.endswith("running_mean")):
            continue   # (running_mean contains that bias.)  The bias gradient is analytically zero (bias feeds BatchNorm); Adam turns its round-off noise into
                       # +-lr steps, so the value is a random walk in the reference too and never affects outputs
        if v.dtype.is_floating_point:
            helpers.assert_close(v, final[k], f"final {k}", tol=1e-4)


def check_lean_collate(device, kind="gcn", compact=None):
    """SubjectStore.collate(prepare_for=...) is LEAN: nothing but what the layer kernels read is written; the reference
    fields and the CSR arrays appear on first access, bit for bit what a full collate writes; a forward-only batch
    (backward=False) still trains - the missing half of the structure is built on demand."""
    from connectome_gnn.graph import ConnectomeBatch, SubjectStore, pack_graphs
    from connectome_gnn.models import GCNConnectome, GraphSAGEConnectome
    from connectome_gnn.synthetic import generate_dataset
    from connectome_gnn.train import CrossEntropyLoss
    graphs = generate_dataset(num_subjects=6, num_regions=84, seed=41) + generate_dataset(num_subjects=2, num_regions=30, seed=42)
    store = SubjectStore(pack_graphs(graphs, compact=compact), device)
    ids = np.array([7, 0, 3, 3, 5, 1])
    full = store.collate(ids)
    lean = store.collate(ids, prepare_for=kind)
    fwd_only = store.collate(ids, prepare_for=kind, backward=False)
    peek = lambda b, f: object.__getattribute__(b, f)
    for b in (lean, fwd_only):
        assert all(peek(b, f) is None for f in ConnectomeBatch._LAZY) and not b.csr.is_full()
        assert all(b.csr.peek(f) is None for f in b.csr._ARRAYS)
    assert lean.csr.agg[kind][1] is not None and fwd_only.csr.agg[kind][1] is None
    torch.manual_seed(0)
    cls = GCNConnectome if kind == "gcn" else GraphSAGEConnectome
    m = cls(in_channels=5, hidden_dim=64, num_classes=2, num_layers=3, dropout=0.0).to(device)
    grads = []
    for b in (full, lean, fwd_only):
        m.zero_grad()
        m.train()
        CrossEntropyLoss()(m(b), b.labels).backward()
        grads.append(torch.cat([p.grad.reshape(-1) for p in m.parameters()]).clone())
        for bn in m.batch_norms:
            bn.reset_running_stats()
    assert torch.equal(grads[0], grads[1]) and torch.equal(grads[0], grads[2])
    if torch.device(device).type == "cuda":   # (the simulator has no tensor-core kernels: its generic ones ask for the arrays)
        assert not lean.csr.is_full(), "a training step on a lean batch must not need the CSR arrays"
    for f in ("node_features", "edge_index", "edge_weight", "batch", "labels", "ptr"):
        assert torch.equal(getattr(lean, f), getattr(full, f)), f
    for f in lean.csr._ARRAYS + ("graph_meta", "eptr"):      # first read adopts the arrays of the same full collate
        assert torch.equal(getattr(lean.csr, f), getattr(full.csr, f)), f
    assert lean.csr.is_full()


def check_fused_eval(device, layers=3, sizes=(84, 84, 360, 30, 84, 57, 84, 84, 130)):
    """cgnn_eval_fused_fwd (the whole eval-mode network in one kernel) against the layer-by-layer path: logits and
    embeddings within 1e-6 max-norm relative, and bit-identical for every subject however the batch is split."""
    from connectome_gnn import _lib
    from connectome_gnn.graph import SubjectStore, pack_graphs
    from connectome_gnn.models import GCNConnectome
    from connectome_gnn.synthetic import generate_connectome
    graphs = [generate_connectome(num_regions=n, seed=100 + s) for s, n in enumerate(sizes)]
    store = SubjectStore(pack_graphs(graphs), device)
    torch.manual_seed(3)
    m = GCNConnectome(in_channels=5, hidden_dim=64, num_classes=2, num_layers=layers, dropout=0.3).to(device)
    with torch.no_grad():                       # non-trivial BatchNorm statistics
        for bn in m.batch_norms:
            bn.running_mean.normal_(0.0, 0.05)
            bn.running_var.uniform_(0.8, 1.2)
            bn.weight.normal_(1.0, 0.1)
            bn.bias.normal_(0.0, 0.1)
    m.eval()
    m.fused_eval = True
    ids = np.arange(len(graphs))
    lib = _lib.load() if torch.device(device).type == "cuda" else None
    with torch.no_grad():
        batch = store.collate(ids, prepare_for="gcn", backward=False)
        before = lib.cgnn_kernel_launches() if lib else 0
        fused_logits, fused_emb = m(batch), m.encode(batch)
        if lib:   # two launches per call: the weight / affine preparation and the fused kernel
            assert lib.cgnn_kernel_launches() - before == 4
        m.fused_eval = False
        ref_logits, ref_emb = m(store.collate(ids, prepare_for="gcn")), m.encode(store.collate(ids, prepare_for="gcn"))
        m.fused_eval = True
        helpers.assert_close(fused_logits, ref_logits, "fused eval logits vs layer by layer", tol=1e-6)
        helpers.assert_close(fused_emb, ref_emb, "fused eval embeddings vs layer by layer", tol=1e-6)
        # any split, any order: every subject's logits are the same bits
        n_s = len(ids)
        for part in (ids[::-1].copy(), ids[: n_s // 2], ids[n_s // 2:], np.array([2]), np.array([n_s - 1, 2, n_s // 2])):
            out = m(store.collate(part, prepare_for="gcn", backward=False))
            assert torch.equal(out, fused_logits[torch.from_numpy(part).to(out.device)]), part
    # a call that wants gradients takes the differentiable path
    m.zero_grad()
    m(store.collate(ids, prepare_for="gcn")).sum().backward()
    assert all(p.grad is not None for p in m.parameters())


def check_fused_step(device, kind="gcn", steps=4):
    """Trainer.enable_fused_step(): flat parameter / gradient buffers + cgnn_adam_step against the plain Trainer with
    torch.optim.Adam on the same batches (dropout 0): identical losses, parameters equal to fp32 round-off of the update."""
    from connectome_gnn.graph import SubjectStore, pack_graphs
    from connectome_gnn.models import GCNConnectome, GraphSAGEConnectome
    from connectome_gnn.synthetic import generate_dataset
    from connectome_gnn.train import Trainer
    graphs = generate_dataset(num_subjects=24, num_regions=30, seed=9)
    store = SubjectStore(pack_graphs(graphs), device)
    cls = GCNConnectome if kind == "gcn" else GraphSAGEConnectome
    runs = {}
    for fused in (False, True):
        torch.manual_seed(5)
        model = cls(in_channels=5, hidden_dim=64, num_classes=2, num_layers=3, dropout=0.0).to(device)
        opt = torch.optim.Adam(model.parameters(), lr=1e-2, weight_decay=1e-3, foreach=False)
        tr = Trainer(model, opt, device=device)
        if fused:
            tr.enable_fused_step()
            assert tr._adam is not None and all(p.data_ptr() >= tr.flat.param.data_ptr() for p in model.parameters())
        losses = []
        for s in range(steps):
            model.train()
            ids = (np.arange(8) * 3 + s) % 24
            losses.append(float(tr.train_step(store.collate(ids, prepare_for=kind))))
        if fused:     # the gradients live in the flat buffer: param.grad aliases it
            for p in model.parameters():
                off, n, _ = tr.flat.slots[id(p)]
                assert p.grad is not None and p.grad.data_ptr() == tr.flat.grad.data_ptr() + 4 * off
        runs[fused] = (losses, {k: v.detach().clone() for k, v in model.state_dict().items()})
    assert runs[True][0][0] == runs[False][0][0], "first-step loss must be identical (same kernels, same weights)"
    for a, b in zip(runs[True][0], runs[False][0]):
        assert abs(a - b) <= 2e-6 * max(abs(b), 1.0), (runs[True][0], runs[False][0])
    for k, v in runs[False][1].items():
        # GCN conv biases feed BatchNorm: their gradient is pure round-off and Adam normalises round-off to
        # full-size steps, so they are not comparable between two summation orders - and do not influence the output
        noise_driven = kind == "gcn" and ((k.startswith("convs.") and k.endswith(".bias")) or k.endswith("running_mean"))
        if v.dtype.is_floating_point and not noise_driven:     # (running_mean absorbs the conv bias)
            helpers.assert_close(runs[True][1][k], v, f"after {steps} fused steps: {k}", tol=2e-5)


def check_adam_kernel(device):
    """cgnn_adam_step against torch.optim.Adam (single-tensor implementation) on the same gradient sequence."""
    from connectome_gnn import _engine
    eng = _engine.engine_for(torch.zeros(1, device=device)) if torch.device(device).type == "cuda" else helpers.emu_engine()
    g = torch.Generator().manual_seed(3)
    n = 5000
    w0 = torch.randn(n, generator=g)
    p_ref = torch.nn.Parameter(w0.clone().to(device))
    opt = torch.optim.Adam([p_ref], lr=3e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4, foreach=False)
    w, m, v = w0.clone().to(device), torch.zeros(n, device=device), torch.zeros(n, device=device)
    state = torch.zeros(4, dtype=torch.int64, device=device)
    worst, equal = 0.0, 0
    for step in range(6):
        grad = (torch.randn(n, generator=g) * (0.5 ** step)).to(device)
        p_ref.grad = grad.clone()
        opt.step()
        eng.step_tick(state)
        eng.adam_step(w, grad, m, v, 3e-3, 0.9, 0.999, 1e-8, 1e-4, state)
        assert int(state[0]) == step + 1
        worst = max(worst, helpers.max_rel(w, p_ref.detach()))
        equal += int(torch.equal(w, p_ref.detach()))
    assert worst <= 1e-6, worst          # one fp32 ulp of the update at most
    return worst, equal


def check_graphed_step(kind="gcn", batch=16, regions=84):
    """Trainer.capture(): one CUDA graph per step.  With dropout off the replays reproduce the eager fused step bit for
    bit (same kernels, same order); with dropout on every replay draws new masks from the device salt; evaluation replays
    equal the eager evaluation.  Returns (train us per step, eval us per step) measured with CUDA events."""
    from connectome_gnn.graph import SubjectStore, pack_graphs
    from connectome_gnn.models import GCNConnectome, GraphSAGEConnectome
    from connectome_gnn.synthetic import generate_dataset
    from connectome_gnn.train import Trainer
    dev = "cuda"
    graphs = generate_dataset(num_subjects=3 * batch, num_regions=regions, seed=21)
    store = SubjectStore(pack_graphs(graphs), dev)
    cls = GCNConnectome if kind == "gcn" else GraphSAGEConnectome
    id_seq = [(np.arange(batch) * 3 + s) % len(graphs) for s in range(6)]

    def make(dropout):
        torch.manual_seed(11)
        model = cls(in_channels=5, hidden_dim=64, num_classes=2, num_layers=3, dropout=dropout).to(dev)
        return Trainer(model, torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4), device=dev)

    eager = make(0.0).enable_fused_step(seed=1)
    ref_losses = []
    for ids in id_seq:
        eager.model.train()
        ref_losses.append(float(eager.train_step(store.collate(ids, prepare_for=kind))))
    graphed = make(0.0)
    graphed.enable_fused_step(seed=1)
    step = graphed.capture(store, batch, kind, train=True)
    # the capture's warm-up steps trained the model: start both from the same weights again
    fresh = make(0.0)
    with torch.no_grad():
        for p, q in zip(graphed.model.parameters(), fresh.model.parameters()):
            p.copy_(q)
        for (_, b1), (_, b2) in zip(graphed.model.named_buffers(), fresh.model.named_buffers()):
            b1.copy_(b2)
        graphed._adam["m"].zero_(); graphed._adam["v"].zero_(); graphed._state[0] = 0
    got = [float(step(ids)) for ids in id_seq]
    assert got == ref_losses, (got, ref_losses)
    for p, q in zip(graphed.model.parameters(), eager.model.parameters()):
        assert torch.equal(p, q)
    # dropout: new masks on every replay
    drop = make(0.3)
    dstep = drop.capture(store, batch, kind, train=True)
    with torch.no_grad():
        saved = [p.detach().clone() for p in drop.model.parameters()]
    l1 = float(dstep(id_seq[0]))
    with torch.no_grad():
        for p, q in zip(drop.model.parameters(), saved):
            p.copy_(q)
    l2 = float(dstep(id_seq[0]))
    assert np.isfinite(l1) and np.isfinite(l2) and l1 != l2, (l1, l2)
    # evaluation replays
    estep = graphed.capture(store, batch, kind, train=False)
    graphed.model.eval()
    for ids in id_seq[:3]:
        loss_g, corr_g = estep(ids)
        lg, cg = float(loss_g), int(corr_g)
        loss_e, corr_e = graphed.eval_step(store.collate(ids, prepare_for=kind, backward=False))
        assert lg == float(loss_e) and cg == int(corr_e)
    # device time per replay
    out = []
    for fn in (lambda: step(id_seq[0]), lambda: estep(id_seq[0])):
        for _ in range(5):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / 50 * 1e3)
    return tuple(out)


def check_batches_die_by_refcount(device, kind="gcn", num_regions=30):
    """A step leaves no reference cycle behind: every batch (and with it GBs of device memory at bench sizes) is freed
    when its last reference goes, not whenever the cyclic garbage collector next runs - with cycles the caching allocator
    keeps asking the driver for new segments in the middle of timed steps."""
    import gc
    import weakref
    from connectome_gnn.graph import SubjectStore, pack_graphs
    from connectome_gnn.models import GCNConnectome, GraphSAGEConnectome
    from connectome_gnn.synthetic import generate_dataset
    from connectome_gnn.train import Trainer
    graphs = generate_dataset(num_subjects=8, num_regions=num_regions, seed=9)
    store = SubjectStore(pack_graphs(graphs), device)
    cls = GCNConnectome if kind == "gcn" else GraphSAGEConnectome
    model = cls(in_channels=5, hidden_dim=64, num_classes=2, num_layers=3, dropout=0.3).to(device)
    tr = Trainer(model, torch.optim.Adam(model.parameters(), lr=1e-2), device=device).enable_fused_step()
    ids = np.arange(8)
    for mode in ("train", "eval", "untouched"):
        for rep in range(2):
            if rep == 1:
                gc.collect()
                gc.disable()
            try:
                b = store.collate(ids, prepare_for=kind, backward=(mode == "train"))
                alive = [weakref.ref(b.csr), weakref.ref(b.node_features)]
                if mode == "train":
                    model.train()
                    tr.train_step(b)
                elif mode == "eval":
                    model.eval()
                    tr.eval_step(b)
                del b
                if rep == 1:
                    assert all(r() is None for r in alive), f"{kind} {mode}: the batch survived its last reference (cycle)"
            finally:
                gc.enable()


def _blob_words(blob, meta, g):
    """The defined words of subject g's aggregation blob: descriptors, then every row's records [begin, end)."""
    nb, n, eb, m = (int(v) for v in meta[g])
    base = 8 * nb + 2 * (eb + (eb & 1)) + 4 * g
    desc = blob[base:base + 4 * n].reshape(n, 4)
    out = [desc.reshape(-1)]
    rec = blob[base + 4 * n:]
    for i in range(n):
        b, e = int(desc[i, 0]), int(desc[i, 1])
        out.append(rec[2 * b:2 * e])
    return torch.cat(out)


def check_pair_collate(device, sizes=(30, 84, 57, 130), with_bad_edge=True):
    """k_collate_pairs (pair store, lean batch: one sort for both blobs) writes bit for bit the blobs, row_graph and
    graph_meta the general collate kernel writes for the same subjects - also with an out-of-range endpoint in the store."""
    from connectome_gnn.graph import SubjectStore, pack_graphs
    from connectome_gnn.synthetic import generate_dataset
    graphs = []
    for k, n in enumerate(sizes):
        graphs += generate_dataset(num_subjects=2, num_regions=n, seed=100 + k)
    packed = pack_graphs(graphs)
    assert packed["edge_pairs"] == 1, "the synthetic generator emits adjacent reversed pairs: pack_graphs must detect them"
    if with_bad_edge:       # endpoint beyond the subject: both kernels turn the edge into a zero-weight self edge on node 0
        packed["src"] = packed["src"].clone()
        packed["src"][5] = int(packed["src"][5]) | (0x7fff << 16)
    store = SubjectStore(packed, device)
    ids = np.array([3, 0, 7, 7, 2, 5, 1])
    for kind in ("gcn", "sage"):
        for backward in (True, False):
            lean = store.collate(ids, prepare_for=kind, backward=backward)
            full = store.collate(ids, prepare_for=kind, lean=False, backward=backward)
            assert not lean.csr.is_full() and full.csr.is_full()
            meta = full.csr.graph_meta.cpu()
            assert torch.equal(lean.csr.graph_meta.cpu(), meta)
            assert torch.equal(lean.node_features, full.node_features) and torch.equal(lean.labels, full.labels)
            la, fa = lean.csr.agg[kind], full.csr.agg[kind]
            assert torch.equal(la[2], fa[2]), "row_graph"
            assert (la[1] is None) == (not backward)
            for d in range(2 if backward else 1):
                lb, fb = la[d].cpu(), fa[d].cpu()
                for g in range(len(ids)):
                    assert torch.equal(_blob_words(lb, meta, g), _blob_words(fb, meta, g)), (kind, backward, d, g)


def check_pair_collate_fuzz(device, seed=0, rounds=6):
    """Random undirected multigraphs - self-loop pairs, repeated pairs, isolated nodes, one-node and edgeless subjects,
    ragged sizes - through the pair store: k_collate_pairs vs k_collate_graph blob for blob, and the lean batch's lazily
    filled reference fields vs the full collate."""
    from connectome_gnn.graph import ConnectomeGraph, SubjectStore, pack_graphs
    rng = np.random.default_rng(seed)
    for rnd in range(rounds):
        graphs = []
        for s in range(int(rng.integers(3, 9))):
            n = int(rng.choice([1, 2, 5, 17, 40, 84, 130, 200]))
            pairs = int(rng.integers(0, 4 * n + 1)) if n > 1 else int(rng.integers(0, 3))
            u = rng.integers(0, n, pairs)
            v = rng.integers(0, n, pairs)            # u == v happens: a self-loop pair; repeats happen: multi-edges
            w = rng.random(pairs).astype(np.float32) + 0.05
            ei = np.empty((2, 2 * pairs), dtype=np.int64)
            ei[0, 0::2], ei[1, 0::2], ei[0, 1::2], ei[1, 1::2] = u, v, v, u
            ew = np.repeat(w, 2)
            graphs.append(ConnectomeGraph(torch.from_numpy(rng.normal(size=(n, 5)).astype(np.float32)), torch.from_numpy(ei),
                                          torch.from_numpy(ew), torch.tensor(int(rng.integers(0, 2)))))
        packed = pack_graphs(graphs)
        if all(g.num_edges == 0 for g in graphs):
            continue
        assert packed["edge_pairs"] == 1
        store = SubjectStore(packed, device)
        ids = rng.permutation(len(graphs))
        for kind in ("gcn", "sage"):
            lean = store.collate(ids, prepare_for=kind)
            full = store.collate(ids, prepare_for=kind, lean=False)
            meta = full.csr.graph_meta.cpu()
            assert torch.equal(lean.csr.graph_meta.cpu(), meta)
            la, fa = lean.csr.agg[kind], full.csr.agg[kind]
            assert torch.equal(la[2], fa[2])
            for d in range(2):
                lb, fb = la[d].cpu(), fa[d].cpu()
                for g in range(len(ids)):
                    assert torch.equal(_blob_words(lb, meta, g), _blob_words(fb, meta, g)), (rnd, kind, d, g)
            for f in ("node_features", "edge_index", "edge_weight", "batch", "labels", "ptr"):
                assert torch.equal(getattr(lean, f), getattr(full, f)), f


def check_pooled_last_layer(device, sizes=(84, 30, 130, 57, 84, 200, 360, 1, 2)):
    """GCN in eval mode under torch.no_grad(): the last layer runs as cgnn_gcn_layer_fwd_pool (BatchNorm with running
    statistics, ReLU and the mean-pool readout folded into the layer kernel, z_L never written).  Against the same model
    with gradients enabled (layer kernel + cgnn_pool_fwd): embeddings and logits within 1e-6 max-norm relative; and the
    embedding of a subject does not depend on what else is in the batch (bit for bit)."""
    from connectome_gnn import _engine
    from connectome_gnn.graph import SubjectStore, pack_graphs
    from connectome_gnn.models import GCNConnectome
    from connectome_gnn.synthetic import generate_connectome
    graphs = [generate_connectome(num_regions=max(n, 10), seed=400 + k) for k, n in enumerate(sizes)]
    store = SubjectStore(pack_graphs(graphs), device)
    ids = np.arange(len(graphs))
    torch.manual_seed(21)
    m = GCNConnectome(in_channels=5, hidden_dim=64, num_classes=2, num_layers=3, dropout=0.4).to(device).eval()
    m.fused_eval = False            # not the whole-network kernel: the layer path
    with torch.no_grad():
        for bn in m.batch_norms:
            bn.running_mean.uniform_(-0.2, 0.2); bn.running_var.uniform_(0.5, 1.5)
            bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-0.3, 0.3)
    eng = _engine.engine_for(store.x)
    calls = {"n": 0}
    orig = eng.gcn_layer_fwd_pool

    def counted(*a, **k):
        out = orig(*a, **k)
        calls["n"] += out is not None
        return out
    eng.gcn_layer_fwd_pool = counted
    try:
        with torch.no_grad():
            emb_fused = m.encode(store.collate(ids, prepare_for="gcn", backward=False)).clone()
            logits_fused = m(store.collate(ids, prepare_for="gcn", backward=False)).clone()
            # every subject alone and in reverse order: same bits
            rev = m.encode(store.collate(ids[::-1].copy(), prepare_for="gcn", backward=False))
            assert torch.equal(rev.flip(0), emb_fused), "embedding depends on the position in the batch"
            for g in (0, 3, 6):
                alone = m.encode(store.collate(np.array([g]), prepare_for="gcn", backward=False))
                assert torch.equal(alone[0], emb_fused[g]), f"subject {g}: embedding depends on the rest of the batch"
        assert calls["n"] >= 3, "the pooled last-layer kernel was not taken"
        before = calls["n"]
        emb_ref = m.encode(store.collate(ids, prepare_for="gcn")).detach()       # gradients enabled: the unfused path
        logits_ref = m(store.collate(ids, prepare_for="gcn")).detach()
        assert calls["n"] == before, "the pooled kernel must not run when a backward pass can follow"
    finally:
        eng.gcn_layer_fwd_pool = orig
    helpers.assert_close(emb_fused, emb_ref, "pooled last layer: embeddings", tol=1e-6)
    helpers.assert_close(logits_fused, logits_ref, "pooled last layer: logits", tol=1e-6)


def check_multi_subject_units(device, subjects=600, regions=84):
    """Enough small subjects that the warp-specialised kernels pack several of them into one unit (on a GPU the units are
    otherwise sized to cover the SMs): engine layers, the pooled last layer and the fused eval kernel against the generic
    SIMT kernels, and a training step's gradients, to 1e-5 max-norm relative."""
    from connectome_gnn import _engine
    from connectome_gnn.graph import SubjectStore
    from connectome_gnn.models import GCNConnectome
    from connectome_gnn.synthetic_fast import generate_packed
    from connectome_gnn.train import CrossEntropyLoss
    store = SubjectStore(generate_packed(subjects, regions, seed=3, device=device if str(device) != "cpu" else "cpu"), device)
    ids = np.arange(subjects)
    eng = _engine.engine_for(store.x)
    torch.manual_seed(4)
    m = GCNConnectome(in_channels=5, hidden_dim=64, num_classes=2, num_layers=3, dropout=0.0).to(device)
    with torch.no_grad():
        for bn in m.batch_norms:
            bn.running_mean.uniform_(-0.2, 0.2); bn.running_var.uniform_(0.5, 1.5)

    state = {k: v.clone() for k, v in m.state_dict().items()}

    def run(tensor_cores: int, fused):
        assert eng.lib.cgnn_set_option(1, tensor_cores) == 0
        try:
            m.load_state_dict(state)         # (the training step below moves the running statistics)
            m.fused_eval = fused
            m.eval()
            with torch.no_grad():
                logits = m(store.collate(ids, prepare_for="gcn", backward=False)).clone()
            m.train()
            m.zero_grad()
            b = store.collate(ids, prepare_for="gcn")
            CrossEntropyLoss()(m(b), b.labels).backward()
            grads = torch.cat([p.grad.reshape(-1) for n, p in m.named_parameters() if not (n.startswith("convs.") and n.endswith(".bias"))])
            return logits, grads.clone()
        finally:
            eng.lib.cgnn_set_option(1, 1)

    ref_logits, ref_grads = run(0, False)               # generic SIMT kernels
    for fused in (False, True):                         # engine + pooled last layer / whole-network kernel
        logits, grads = run(1, fused)
        helpers.assert_close(logits, ref_logits, f"multi-subject units (fused_eval={fused}): eval logits", tol=1e-5)
        helpers.assert_close(grads, ref_grads, f"multi-subject units (fused_eval={fused}): gradients", tol=1e-5)
