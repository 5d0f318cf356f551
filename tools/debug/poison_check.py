"""Do the eval paths read memory nobody wrote?  Fill the caching allocator's free blocks with NaN / huge values between
runs and compare the logits bit for bit."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "connectome-gnn-suite_b200")]
import numpy as np, torch
from connectome_gnn.graph import SubjectStore, pack_graphs
from connectome_gnn.models import GCNConnectome, GraphSAGEConnectome
from connectome_gnn.synthetic import generate_dataset
dev = "cuda"
pool = generate_dataset(num_subjects=16, num_regions=360, seed=42)
store = SubjectStore(pack_graphs(pool), dev)

def poison(value):
    blocks = [torch.full((n,), value, device=dev) for n in (1 << 26, 1 << 24, 1 << 22, 1 << 20, 1 << 18, 1 << 16, 1 << 14, 1 << 12) for _ in range(3)]
    torch.cuda.synchronize()
    del blocks

for kind, cls in (("sage", GraphSAGEConnectome), ("gcn", GCNConnectome)):
    for mode in (False, True):
        torch.manual_seed(0)
        m = cls(in_channels=5, hidden_dim=64, num_classes=2, num_layers=3, dropout=0.0).to(dev).eval()
        m.fused_eval = mode
        outs = []
        for value in (0.0, float("nan"), 3e38):
            poison(value)
            with torch.no_grad():
                a = m(store.collate(np.arange(16), prepare_for=kind))
                b = m(store.collate(np.array([5])))
                c = m(store.collate(np.arange(3, 9), prepare_for=kind, backward=False))
            outs.append(torch.cat([a, b, c]).clone())
        print(kind, mode, "finite", bool(torch.isfinite(torch.stack(outs)).all()), "equal", bool(torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])), flush=True)
    # training step as well
    m.train()
    from connectome_gnn.train import CrossEntropyLoss
    gs = []
    for value in (0.0, float("nan"), 3e38):
        poison(value)
        m.zero_grad()
        for bn in m.batch_norms: bn.reset_running_stats()
        batch = store.collate(np.arange(16), prepare_for=kind)
        CrossEntropyLoss()(m(batch), batch.labels).backward()
        gs.append(torch.cat([p.grad.reshape(-1) for p in m.parameters()]).clone())
    print(kind, "train grads finite", bool(torch.isfinite(torch.stack(gs)).all()), "equal", bool(torch.equal(gs[0], gs[1]) and torch.equal(gs[0], gs[2])), flush=True)
