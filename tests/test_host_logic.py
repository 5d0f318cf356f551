"""Host-side logic that needs no GPU: subject packing, batch planning / sharding, Trainer control flow."""
import math

import numpy as np
import pytest
import torch

import helpers


def _graph(n, e, f=3, label=0, seed=0):
    from connectome_gnn.graph import ConnectomeGraph
    g = torch.Generator().manual_seed(seed)
    src, dst = torch.randint(0, n, (e,), generator=g), torch.randint(0, n, (e,), generator=g)
    return ConnectomeGraph(torch.randn(n, f, generator=g), torch.stack([src, dst]), torch.rand(e, generator=g),
                           None if label is None else torch.tensor(label), f"s{seed}")


def test_connectome_graph_helpers():
    g = _graph(10, 25)
    assert (g.num_nodes, g.num_edges, g.num_features) == (10, 25, 3)
    A = g.adjacency_matrix()
    assert A.shape == (10, 10) and float(A[g.edge_index[0, -1], g.edge_index[1, -1]]) == float(g.edge_weight[-1])
    deg = g.degree()
    assert deg.shape == (10,) and torch.allclose(deg.sum(), g.edge_weight.sum())
    assert g.to("cpu").node_features.device.type == "cpu"


def test_pack_graphs_layout_and_validation():
    from connectome_gnn.graph import pack_graphs
    graphs = [_graph(5, 7, seed=1), _graph(9, 0, seed=2, label=None), _graph(4, 12, seed=3, label=1)]
    p = pack_graphs(graphs, compact=False)
    assert p["node_ptr"].tolist() == [0, 5, 14, 18] and p["edge_ptr"].tolist() == [0, 7, 7, 19]
    assert p["src"].dtype == torch.int32 and p["x"].shape == (18, 3) and p["num_features"] == 3
    assert p["has_label"].tolist() == [True, False, True] and p["label"].tolist() == [0, 0, 1]
    assert torch.equal(p["src"][7:].long(), graphs[2].edge_index[0])
    c = pack_graphs(graphs)      # default: the smallest form the lists allow - compact words; these random lists are not pair lists
    assert c["edge_pairs"] == 0 and c["dst"].numel() == 0
    assert torch.equal(c["src"][7:].long() & 0xffff, graphs[2].edge_index[0]) and torch.equal((c["src"][7:].long() >> 16) & 0xffff, graphs[2].edge_index[1])
    bad = _graph(5, 7, seed=4)
    bad.edge_index[1, 3] = 5
    with pytest.raises(ValueError, match="outside its subject"):
        pack_graphs([bad])
    with pytest.raises(ValueError, match="same number of node features"):
        pack_graphs([_graph(5, 7, f=3), _graph(5, 7, f=4)])
    with pytest.raises(ValueError):
        pack_graphs([])


@pytest.mark.parametrize("count,world", [(16, 8), (13, 4), (3, 8), (0, 2), (4096, 8), (7, 1)])
def test_shard_bounds_partition_a_chunk(count, world):
    from connectome_gnn.graph import shard_bounds
    spans = [shard_bounds(count, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == count
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def test_loader_len_order_and_plan():
    from connectome_gnn.graph import ConnectomeDataLoader
    ds = [_graph(4 + i % 3, 6, seed=i) for i in range(23)]
    loader = ConnectomeDataLoader(ds, batch_size=8, shuffle=False, rank=0, world_size=1)
    assert len(loader) == math.ceil(23 / 8) == 3
    assert loader.epoch_order() == list(range(23))
    plan = loader.plan(loader.epoch_order())
    assert [len(s["ids"]) for s in plan] == [8, 8, 7] and [s["global_num_graphs"] for s in plan] == [8, 8, 7]
    # shuffling draws torch.randperm from the global generator, exactly like the reference loader (graph.py:193)
    shuffled = ConnectomeDataLoader(ds, batch_size=8, shuffle=True, rank=0, world_size=1)
    torch.manual_seed(5)
    expect = torch.randperm(23).tolist()
    torch.manual_seed(5)
    assert shuffled.epoch_order() == expect
    # two ranks: same permutation, complementary contiguous slices of every chunk
    parts = []
    for r in range(2):
        l = ConnectomeDataLoader(ds, batch_size=8, shuffle=False, rank=r, world_size=2)
        parts.append(l.plan(l.epoch_order()))
    for s0, s1, whole in zip(parts[0], parts[1], plan):
        assert np.concatenate([s0["ids"], s1["ids"]]).tolist() == whole["ids"].tolist()
        assert s0["graph_base"] == 0 and s1["graph_base"] == len(s0["ids"])


class _ScriptedTrainer:
    """Trainer.fit's control flow with scripted epoch results (no model arithmetic)."""

    def __new__(cls, val_losses):
        from connectome_gnn.train import Trainer

        class T(Trainer):
            def __init__(self, losses):
                self.model = torch.nn.Linear(1, 1)
                self.losses, self.epoch, self.snapshots = list(losses), 0, []

            def train_epoch(self, loader):
                self.epoch += 1
                with torch.no_grad():
                    self.model.weight.fill_(float(self.epoch))
                return 1.0 / self.epoch

            def evaluate(self, loader):
                return {"loss": self.losses[self.epoch - 1], "accuracy": 0.5, "correct": 1, "total": 2}

        return T(val_losses)


def test_fit_early_stopping_and_best_state_restore():
    # best at epoch 2; patience 3 -> stops after epoch 5 (reference train.py:113-121: strict <, epoch - best >= patience)
    t = _ScriptedTrainer([0.9, 0.5, 0.5, 0.6, 0.7, 0.1, 0.1])
    hist = t.fit(None, None, num_epochs=7, patience=3, verbose=False)
    assert len(hist["train_loss"]) == 5 and hist["val_loss"] == [0.9, 0.5, 0.5, 0.6, 0.7]
    assert float(t.model.weight) == 2.0, "weights of the best epoch (2) are restored; the tie at epoch 3 does not count"
    # no early stop: runs to completion and still restores the best epoch (train.py:124-125)
    t = _ScriptedTrainer([0.9, 0.8, 0.3, 0.4])
    hist = t.fit(None, None, num_epochs=4, patience=10, verbose=False)
    assert len(hist["val_acc"]) == 4 and float(t.model.weight) == 3.0
    assert set(hist) == {"train_loss", "val_loss", "val_acc"}


def test_state_dict_keys_match_reference():
    """checkpoints must move between the reference and this package unchanged."""
    from connectome_gnn.models import GCNConnectome, GraphSAGEConnectome
    a = helpers.golden("ref_small.npz")
    for kind, cls in (("gcn", GCNConnectome), ("sage", GraphSAGEConnectome)):
        ref = helpers.state_dict_from(a, f"{kind}.init")
        m = cls(in_channels=5, hidden_dim=16, num_classes=2, num_layers=3, dropout=0.1)
        sd = m.state_dict()
        assert list(sd) == list(ref)
        assert all(sd[k].shape == ref[k].shape and sd[k].dtype == ref[k].dtype for k in ref)
        m.load_state_dict(ref)
        assert m.dropout == 0.1 and len(m.convs) == len(m.batch_norms) == 3 and len(m.classifier) == 4
    assert sum(p.numel() for p in GCNConnectome(5, 64).parameters()) == 11234
    assert sum(p.numel() for p in GraphSAGEConnectome(5, 64).parameters()) == 19746


def test_generator_properties():
    """Mirrors the reference's tests/test_synthetic.py assertions on the package's generator."""
    from connectome_gnn import REGION_NAMES, generate_connectome, generate_dataset
    from connectome_gnn.synthetic import small_world_stats
    g = generate_connectome(num_regions=84, seed=0)
    assert g.num_nodes == 84 and g.num_features == 5 and g.num_edges == 8 * 84
    assert (g.edge_weight > 0).all() and int(g.label) in (0, 1)
    pairs = set(zip(g.edge_index[0].tolist(), g.edge_index[1].tolist()))
    assert all((v, u) in pairs for u, v in pairs)
    g2 = generate_connectome(num_regions=84, seed=0)
    assert torch.equal(g.edge_index, g2.edge_index) and torch.equal(g.node_features, g2.node_features)
    assert not torch.equal(generate_connectome(seed=1).node_features, generate_connectome(seed=999).node_features)
    ds = generate_dataset(num_subjects=60, num_regions=20, seed=0)
    assert 5 < sum(int(x.label) for x in ds) < 55 and ds[3].subject_id == "sub-0003"
    stats = small_world_stats(ds[:5])
    assert stats["num_graphs"] == 5 and stats["mean_clustering"] >= 0 and stats["mean_avg_path_length"] > 1
    assert len(REGION_NAMES) == 83


def test_peer_exchange_slot_schedule():
    """PeerExchange alternates two physical slots per logical exchange and numbers the uses from 1 (a zeroed flag never
    matches): the property the kernel's reuse argument rests on.  Pure host logic - no device, no process group."""
    from connectome_gnn.peers import PeerExchange, _SLOTS
    px = PeerExchange.__new__(PeerExchange)
    px.seq = {}
    seen = {}
    for use in range(1, 7):
        for logical in (0, 1, 5):
            slot, seq = px._slot(logical)
            assert seq == use and slot == 2 * logical + (use & 1) and 0 <= slot < _SLOTS
            assert seen.get(logical) != slot, "two consecutive uses of one exchange must not share a physical slot"
            seen[logical] = slot
    with pytest.raises(RuntimeError):
        px._slot(_SLOTS // 2)
