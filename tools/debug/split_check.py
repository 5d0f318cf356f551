import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "connectome-gnn-suite_b200")]
import numpy as np, torch
from connectome_gnn.graph import SubjectStore, pack_graphs
from connectome_gnn.models import GCNConnectome, GraphSAGEConnectome
from connectome_gnn.synthetic import generate_dataset
dev = "cuda"
B, U = 4096, 64
pool = generate_dataset(num_subjects=U, num_regions=360, seed=42)
graphs = [pool[i % U] for i in range(B)]
store = SubjectStore(pack_graphs(graphs), dev)
ids = np.random.default_rng(2).permutation(B)
for kind, cls in (("gcn", GCNConnectome), ("sage", GraphSAGEConnectome), ("gcn", GCNConnectome)):
    torch.manual_seed(0)
    m = cls(in_channels=5, hidden_dim=64, num_classes=2, num_layers=3, dropout=0.3).to(dev).eval()
    for mode in ("auto", False):
        m.fused_eval = mode
        with torch.no_grad():
            full = m(store.collate(ids, prepare_for=kind))
            full2 = m(store.collate(ids, prepare_for=kind))
            full3 = m(store.collate(ids))
            parts = torch.cat([m(store.collate(ids[lo:hi])) for lo, hi in ((0, 1000), (1000, 1001), (1001, B))])
        for name, other in (("same call twice", full2), ("full collate", full3), ("parts", parts)):
            d = (full - other).abs()
            bad = (d.max(dim=1).values > 0).nonzero().flatten()
            print(kind, mode, name, "max diff", float(d.max()), "rows differing", bad.numel(), bad[:10].tolist(), flush=True)
