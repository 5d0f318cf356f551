"""Data-parallel host logic under torch.distributed (gloo, world_size 2, CPU): subject sharding, SyncBN statistic
merge, BatchNorm-backward sum reduction, the flat gradient all-reduce and the evaluate() totals must reproduce the
single-process global batch.  The kernels run on the test-only simulator (tests/emu), so this exercises the
same C entry points and the same Python plumbing the NCCL path uses, without a GPU."""
import os
import socket
import sys

import pytest
import torch

import helpers

ROOT = helpers.ROOT


def _setup_emu():
    from connectome_gnn import _engine
    eng = helpers.emu_engine()
    _engine.engine_for = lambda t: eng
    _engine.default_device = lambda: torch.device("cpu")
    return eng


def _run_case(kind, graphs, batch_size, rank, world):
    """One optimizer step on the first batch + evaluate on everything; returns (loss share, state_dict, eval dict)."""
    from connectome_gnn.graph import ConnectomeDataLoader
    from connectome_gnn.models import GCNConnectome, GraphSAGEConnectome
    from connectome_gnn.train import Trainer
    torch.manual_seed(0)
    cls = GCNConnectome if kind == "gcn" else GraphSAGEConnectome
    model = cls(in_channels=5, hidden_dim=16, num_classes=2, num_layers=2, dropout=0.25)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    trainer = Trainer(model, opt, device="cpu")
    loader = ConnectomeDataLoader(graphs, batch_size=batch_size, shuffle=False, rank=rank, world_size=world)
    batch = next(iter(loader))
    torch.manual_seed(11)            # same dropout stream seed on every rank and in the single-process run
    model.train()
    loss = trainer.train_step(batch)
    ev = trainer.evaluate(ConnectomeDataLoader(graphs, batch_size=3, shuffle=False, rank=rank, world_size=world))
    return float(loss), {k: v.clone() for k, v in model.state_dict().items()}, ev


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "connectome-gnn-suite_b200"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _setup_emu()
        from connectome_gnn.synthetic import generate_dataset
        graphs = generate_dataset(num_subjects=7, num_regions=20, seed=5)
        res = {}
        for kind in ("gcn", "sage"):
            res[kind] = _run_case(kind, graphs, 7, rank, world)          # 4 + 3 subjects
            res[kind + "_tiny"] = _run_case(kind, graphs, 1, rank, world)   # rank 1 holds nothing in the first batch
        torch.save(res, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(600)
def test_two_ranks_reproduce_the_single_process_batch(tmp_path):
    import torch.multiprocessing as mp
    from connectome_gnn.synthetic import generate_dataset
    graphs = generate_dataset(num_subjects=7, num_regions=20, seed=5)
    # single-process expectation (no process group in this process)
    import connectome_gnn._engine as _engine
    saved = (_engine.engine_for, _engine.default_device)
    try:
        _setup_emu()
        expect = {}
        for kind in ("gcn", "sage"):
            expect[kind] = _run_case(kind, graphs, 7, 0, 1)
            expect[kind + "_tiny"] = _run_case(kind, graphs, 1, 0, 1)
    finally:
        _engine.engine_for, _engine.default_device = saved
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    ranks = [torch.load(os.path.join(tmp_path, f"rank{r}.pt"), weights_only=False) for r in range(2)]
    for case, (loss, state, ev) in expect.items():
        shares = [ranks[r][case][0] for r in range(2)]
        assert sum(shares) == pytest.approx(loss, rel=1e-5), (case, shares, loss)
        for r in range(2):
            got_state, got_ev = ranks[r][case][1], ranks[r][case][2]
            for k, v in state.items():
                if v.dtype.is_floating_point:
                    helpers.assert_close(got_state[k], v, f"{case} rank {r}: {k}", tol=2e-5, atol=1e-6)
                else:
                    assert torch.equal(got_state[k], v), (case, r, k)
            assert got_ev["total"] == ev["total"] and got_ev["correct"] == ev["correct"], (case, r, got_ev, ev)
            assert got_ev["loss"] == pytest.approx(ev["loss"], rel=1e-5)
            assert got_ev["accuracy"] == pytest.approx(ev["accuracy"])
