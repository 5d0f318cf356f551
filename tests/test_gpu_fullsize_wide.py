"""GPU parity at BASELINE.json configs[4]'s per-GPU size: GCN, 360-node subjects, hidden 256, 8192 subjects per GPU
(2.95 M rows; one activation tensor is 3 GB, so element offsets pass 2^31 bytes).

Same plan as test_gpu_fullsize.py: size-independent properties at the full size (evaluation logits bit-identical
under any batch split and for identical subjects; a permuted training batch gives the same loss and BatchNorm
statistics within 1e-5 and the same gradients within 1e-4 - see the note there) and a direct comparison with the oracle in fp64 at a size it finishes in seconds."""
import numpy as np
import pytest
import torch

import helpers
from helpers import REL_TOL

pytestmark = pytest.mark.gpu
DEV = "cuda"
UNIQUE, BATCH, REGIONS, HIDDEN, LAYERS = 64, 8192, 360, 256, 3


@pytest.fixture(scope="module")
def dataset():
    from connectome_gnn.graph import SubjectStore, pack_graphs
    from connectome_gnn.synthetic import generate_dataset
    pool = generate_dataset(num_subjects=UNIQUE, num_regions=REGIONS, k=8, beta=0.15, trait_idx=0, seed=43)
    graphs = (pool * (BATCH // UNIQUE))[:BATCH]
    return pool, graphs, SubjectStore(pack_graphs(graphs, compact=True, pairs=True), DEV)


def _model(dropout=0.0, seed=0):
    from connectome_gnn.models import GCNConnectome
    torch.manual_seed(seed)
    return GCNConnectome(in_channels=5, hidden_dim=HIDDEN, num_classes=2, num_layers=LAYERS, dropout=dropout).to(DEV)


def test_eval_logits_do_not_depend_on_the_batch_split(dataset):
    pool, graphs, store = dataset
    model = _model().eval()
    ids = np.random.default_rng(2).permutation(BATCH)
    with torch.no_grad():
        full = model(store.collate(ids, prepare_for="gcn"))
        parts = torch.cat([model(store.collate(ids[lo:hi])) for lo, hi in ((0, 3000), (3000, 3001), (3001, BATCH))])
    assert torch.isfinite(full).all()
    assert torch.equal(full, parts), "eval-mode logits changed with the batch split"
    sub = torch.from_numpy(ids % UNIQUE).to(DEV)
    rep = torch.zeros(UNIQUE, 2, device=DEV).index_copy_(0, sub, full)
    assert torch.equal(rep[sub], full), "identical subjects produced different logits"


def test_training_step_is_invariant_to_batch_order(dataset):
    from connectome_gnn.train import CrossEntropyLoss
    pool, graphs, store = dataset
    model = _model().train()
    perm = np.random.default_rng(3).permutation(BATCH)
    out = []
    for order in (np.arange(BATCH), perm):
        for bn in model.batch_norms:
            bn.reset_running_stats()
        model.zero_grad()
        batch = store.collate(order, prepare_for="gcn")
        loss = CrossEntropyLoss()(model(batch), batch.labels)
        loss.backward()
        out.append((float(loss.detach()), torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone(),
                    torch.cat([bn.running_var for bn in model.batch_norms]).clone()))
        del batch, loss
    assert out[0][0] == pytest.approx(out[1][0], rel=1e-6)
    off = 0
    for name, p in model.named_parameters():
        a, b = out[0][1][off:off + p.numel()], out[1][1][off:off + p.numel()]
        off += p.numel()
        print(f"{name:32s} |g| {float(a.abs().max()):.3e}  permutation diff {float((a - b).abs().max()):.3e}")
    # 2.95 M rows: every weight gradient is a sum of 2.95 M terms of both signs, and a permutation changes the order of
    # the fp32 chains inside a CTA (dW lives in tensor memory over ~20 k rows per CTA).  Measured: 5e-5 of the gradient
    # scale; the fp32 reference's own order sensitivity at this subject size is 3e-3 and its distance from
    # the fp64 values is 8.5e-4 where this library's is 3.5e-5 (test below).  The bar here is therefore 1e-4.
    helpers.assert_close(out[1][1], out[0][1], "gradients under a batch permutation", tol=1e-4)
    helpers.assert_close(out[1][2], out[0][2], "BatchNorm running variance under a batch permutation", tol=REL_TOL)


def test_training_step_against_the_oracle_in_fp64(dataset):
    import parity
    from connectome_gnn.train import CrossEntropyLoss
    pool, graphs, store = dataset
    n = 96
    model = _model(seed=5)
    params = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    refs = parity.oracle_reference("gcn", params, graphs[:n])
    f64, f32 = refs["f64"], refs["f32"]
    model.train()
    batch = store.collate(np.arange(n), prepare_for="gcn")
    logits = model(batch)
    loss = CrossEntropyLoss()(logits, batch.labels)
    loss.backward()
    helpers.assert_close(logits, f64["gcn.train.logits"], f"logits, {n} x 360-node subjects, hidden 256", tol=REL_TOL)
    assert float(loss.detach()) == pytest.approx(float(f64["gcn.train.loss"]), rel=1e-5)
    names = [k for k, _ in model.named_parameters()]
    cat = lambda src: torch.cat([torch.as_tensor(src[f"gcn.train.grad.{k}"]).reshape(-1).double() for k in names])
    ours = helpers.max_rel(torch.cat([p.grad.reshape(-1) for _, p in model.named_parameters()]), cat(f64))
    theirs = helpers.max_rel(cat(f32), cat(f64))
    print(f"hidden 256: gradient distance from the fp64 oracle: this library {ours:.2e}, fp32 oracle {theirs:.2e}")
    assert ours <= max(REL_TOL, 2 * theirs), (ours, theirs)
    for k, v in model.state_dict().items():
        if "running" in k:
            helpers.assert_close(v, f64[f"gcn.train.after.{k}"], f"gcn {k}", tol=REL_TOL)
