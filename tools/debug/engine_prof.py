"""A handful of launches of the warp-specialised forward layer at the bench shape, for ncu.
usage: python tools/debug/engine_prof.py [subjects] [regions] [kind]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "connectome-gnn-suite_b200")]
import numpy as np
import torch
from connectome_gnn import _engine
from connectome_gnn._engine import Act
from connectome_gnn.graph import SubjectStore, pack_graphs
from connectome_gnn.synthetic import generate_dataset

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 360
kind = sys.argv[3] if len(sys.argv) > 3 else "gcn"
dev = torch.device("cuda", 0)
pool = generate_dataset(num_subjects=min(B, 64), num_regions=N, seed=42)
graphs = (pool * (-(-B // len(pool))))[:B]
store = SubjectStore(pack_graphs(graphs), dev)
batch = store.collate(np.arange(B), prepare_for=kind)
eng = _engine.engine_for(batch.node_features)
rows = batch.num_nodes
torch.manual_seed(0)
t_in = torch.randn(rows, 64, device=dev)
W = torch.randn(64, 128 if kind == "sage" else 64, device=dev) * 0.2
bias = torch.randn(64, device=dev) * 0.1
scale, shift = torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev) * 0.1
act = Act(scale, shift, kind == "gcn", 0.3, 1234, 1, 0)
for want in (True, False, True, False):
    eng.layer_fwd(kind, t_in, act, W, bias, batch.csr, batch.ptr, B, want_stats=want)
torch.cuda.synchronize()
print("done")
