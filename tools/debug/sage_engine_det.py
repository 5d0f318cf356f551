import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "connectome-gnn-suite_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from connectome_gnn.graph import SubjectStore, pack_graphs
from connectome_gnn.models import GraphSAGEConnectome, GCNConnectome
from connectome_gnn.synthetic import generate_connectome
sizes = [84] * 16
graphs = [generate_connectome(num_regions=n, seed=300 + k) for k, n in enumerate(sizes)]
store = SubjectStore(pack_graphs(graphs), "cuda")
torch.manual_seed(0)
kind = os.environ.get("KIND", "sage")
m = (GraphSAGEConnectome if kind == "sage" else GCNConnectome)(in_channels=5, hidden_dim=64, num_classes=2, num_layers=2, dropout=0.25).to("cuda").eval()
ids = np.arange(len(sizes))
m.fused_eval = False
outs = []
for _ in range(12):
    with torch.no_grad():
        outs.append(m.encode(store.collate(ids, prepare_for=kind, backward=False)).clone())
ref = torch.stack(outs).median(0).values
print("ENG_DBG", kind, os.environ.get("ENG_DBG"), ["%.1e" % float((o - ref).abs().max() / ref.abs().max()) for o in outs])
