"""Same launch, same bits: repeats the GCN / GraphSAGE forward (engine, fused eval, gather + contraction) and a training step's
gradients many times and compares bitwise.  (How the N = 128 tcgen05.mma hazard of DESIGN section 9 was caught.)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "connectome-gnn-suite_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from connectome_gnn.graph import SubjectStore, pack_graphs
from connectome_gnn.models import GraphSAGEConnectome, GCNConnectome
from connectome_gnn.synthetic import generate_connectome
from connectome_gnn.train import CrossEntropyLoss

REPS = int(os.environ.get("REPS", 60))
bad = 0
for sizes in ([84] * 16, [360] * 12, [84, 30, 130, 57, 84, 200, 360]):
    graphs = [generate_connectome(num_regions=n, seed=300 + k) for k, n in enumerate(sizes)]
    store = SubjectStore(pack_graphs(graphs), "cuda")
    ids = np.arange(len(sizes))
    for kind, cls in (("gcn", GCNConnectome), ("sage", GraphSAGEConnectome)):
        for fused in (("auto", False) if kind == "gcn" else (False,)):
            torch.manual_seed(0)
            m = cls(in_channels=5, hidden_dim=64, num_classes=2, num_layers=3, dropout=0.25).to("cuda")
            m.fused_eval = fused
            m.eval()
            outs = []
            for _ in range(REPS):
                with torch.no_grad():
                    outs.append(m(store.collate(ids, prepare_for=kind, backward=False)).clone())
            n_eval = sum(not torch.equal(o, outs[0]) for o in outs)
            m.train()
            grads = []
            for _ in range(REPS // 3):
                m.zero_grad()
                for bn in m.batch_norms:
                    bn.reset_running_stats()
                torch.manual_seed(5)
                b = store.collate(ids, prepare_for=kind)
                CrossEntropyLoss()(m(b), b.labels).backward()
                grads.append(torch.cat([p.grad.reshape(-1) for p in m.parameters()]).clone())
            n_train = sum(not torch.equal(g, grads[0]) for g in grads)
            bad += n_eval + n_train
            print(f"sizes={len(sizes)}x{sizes[0]} {kind} fused_eval={fused}: eval differing {n_eval}/{REPS}, train grads differing {n_train}/{REPS // 3}", flush=True)
print("DETERMINISTIC" if bad == 0 else f"NONDETERMINISTIC: {bad}")
