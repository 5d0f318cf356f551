"""Data-parallel path on the real kernels: two ranks (two processes sharing cuda:0) must reproduce the single-process
global batch - SyncBN statistics, BatchNorm-backward sums, the flat gradient all-reduce, sharding-independent dropout
masks, evaluate() totals.  NCCL refuses two ranks on one device, so the process group is gloo; gloo has no
CUDA all_gather, hence the one all_gather_into_tensor call of the package is staged through host memory HERE (test
plumbing only - under torchrun the package talks NCCL directly, see bench.py --gpus N)."""
import os
import socket
import sys

import pytest
import torch

import helpers

pytestmark = pytest.mark.gpu
ROOT = helpers.ROOT
CASES = (("gcn", 64, 84), ("sage", 64, 84), ("gcn", 256, 45), ("sage", 256, 45))


def _run_case(kind, hidden, graphs, batch_size, rank, world):
    from connectome_gnn.graph import ConnectomeDataLoader
    from connectome_gnn.models import GCNConnectome, GraphSAGEConnectome
    from connectome_gnn.train import Trainer
    torch.manual_seed(0)
    cls = GCNConnectome if kind == "gcn" else GraphSAGEConnectome
    model = cls(in_channels=5, hidden_dim=hidden, num_classes=2, num_layers=3, dropout=0.25)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    trainer = Trainer(model, opt, device="cuda")
    loader = ConnectomeDataLoader(graphs, batch_size=batch_size, shuffle=False, rank=rank, world_size=world)
    batch = next(iter(loader))
    torch.manual_seed(11)            # same dropout stream seed on every rank and in the single-process run
    model.train()
    loss = trainer.train_step(batch)
    ev = trainer.evaluate(ConnectomeDataLoader(graphs, batch_size=4, shuffle=False, rank=rank, world_size=world))
    return float(loss), {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}, ev


def _graphs(regions):
    from connectome_gnn.synthetic import generate_dataset
    return generate_dataset(num_subjects=9, num_regions=regions, seed=5)


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "connectome-gnn-suite_b200"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orig = dist.all_gather_into_tensor

    def staged(out, inp, group=None, async_op=False):
        if not inp.is_cuda:
            return orig(out, inp, group=group, async_op=async_op)
        host = torch.empty(out.shape, dtype=out.dtype)
        orig(host, inp.cpu(), group=group)
        out.copy_(host)
        return None

    dist.all_gather_into_tensor = staged
    try:
        torch.cuda.set_device(0)
        res = {}
        for kind, hidden, regions in CASES:
            graphs = _graphs(regions)
            res[(kind, hidden)] = _run_case(kind, hidden, graphs, 9, rank, world)           # 5 + 4 subjects
            res[(kind, hidden, "tiny")] = _run_case(kind, hidden, graphs, 1, rank, world)   # rank 1 holds nothing
        torch.save(res, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(900)
def test_two_ranks_on_the_device_reproduce_the_single_process_batch(tmp_path):
    import torch.multiprocessing as mp
    expect = {}
    for kind, hidden, regions in CASES:
        graphs = _graphs(regions)
        expect[(kind, hidden)] = _run_case(kind, hidden, graphs, 9, 0, 1)
        expect[(kind, hidden, "tiny")] = _run_case(kind, hidden, graphs, 1, 0, 1)
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
