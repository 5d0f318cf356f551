import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "connectome-gnn-suite_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from connectome_gnn import _engine
from connectome_gnn._engine import Act
from connectome_gnn.graph import SubjectStore, pack_graphs
from connectome_gnn.synthetic import generate_connectome
sizes = [int(s) for s in os.environ.get("SIZES", "84,84,84,84").split(",")]
graphs = [generate_connectome(num_regions=n, seed=300 + k) for k, n in enumerate(sizes)]
store = SubjectStore(pack_graphs(graphs), "cuda")
ids = np.arange(len(sizes))
b = store.collate(ids, prepare_for="sage")
eng = _engine.engine_for(b.node_features)
rows = b.num_nodes
torch.manual_seed(1)
t = torch.randn(rows, 64, device="cuda")
W = torch.randn(64, 128, device="cuda") * 0.1
bias = torch.randn(64, device="cuda") * 0.1
act = Act()
zr, _, agg = eng.layer_fwd("sage", t, act, W, bias, b.csr, b.ptr, len(sizes), want_stats=False, need_agg=True)
for rep in range(int(os.environ.get('REPS', 8))):
    z, _, agg2 = eng.layer_fwd("sage", t, act, W, bias, b.csr, b.ptr, len(sizes), want_stats=False, need_agg=False)
    assert agg2 is None
    d = (z - zr).abs()
    bad = (d > 1e-4).nonzero()
    if bad.numel() == 0:
        pass
    else:
        r = bad[:, 0].unique().tolist(); c = bad[:, 1].unique().tolist()
        print(rep, "bad rows", len(r), r[:24], "cols", c[:40], "max", float(d.max()))
print("done", os.environ.get("ENG_DBG"))
