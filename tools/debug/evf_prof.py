"""A few fused eval forwards (cgnn_eval_fused_fwd) at the bench shape, timed with CUDA events; also the ncu target.
usage: python tools/debug/evf_prof.py [subjects] [regions] [layers]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "connectome-gnn-suite_b200")]
import numpy as np
import torch
from connectome_gnn.graph import SubjectStore, pack_graphs
from connectome_gnn.models import GCNConnectome
from connectome_gnn.synthetic import generate_dataset

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 360
L = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
pool = generate_dataset(num_subjects=min(B, 64), num_regions=N, seed=42)
graphs = (pool * (-(-B // len(pool))))[:B]
store = SubjectStore(pack_graphs(graphs), dev)
batch = store.collate(np.arange(B), prepare_for="gcn", backward=False)
torch.manual_seed(0)
m = GCNConnectome(in_channels=5, hidden_dim=64, num_classes=2, num_layers=L, dropout=0.3).to(dev).eval()
with torch.no_grad():
    for fused in (True, False):
        m.fused_eval = fused
        for _ in range(2):
            out = m(batch)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out = m(batch)
        e1.record()
        torch.cuda.synchronize()
        print(f"fused={fused}: {e0.elapsed_time(e1) / 5 * 1e3:.1f} us per forward of {B} x {N}-node subjects, {L} layers; logits[0] {out[0].tolist()}", flush=True)
