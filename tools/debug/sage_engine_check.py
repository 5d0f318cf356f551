"""Forward-only GraphSAGE (engine, projection first) against the gather + contraction path, over a few shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "connectome-gnn-suite_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from connectome_gnn.graph import SubjectStore, pack_graphs
from connectome_gnn.models import GraphSAGEConnectome
from connectome_gnn.synthetic import generate_connectome

def run(sizes, rand_bn, layers=3, reps=3):
    graphs = [generate_connectome(num_regions=n, seed=300 + k) for k, n in enumerate(sizes)]
    store = SubjectStore(pack_graphs(graphs), "cuda")
    torch.manual_seed(0)
    m = GraphSAGEConnectome(in_channels=5, hidden_dim=64, num_classes=2, num_layers=layers, dropout=0.25).to("cuda").eval()
    if rand_bn:
        with torch.no_grad():
            for bn in m.batch_norms:
                bn.running_mean.uniform_(-0.2, 0.2); bn.running_var.uniform_(0.5, 1.5)
                bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-0.3, 0.3)
    ids = np.arange(len(sizes))
    r = m.encode(store.collate(ids, prepare_for="sage")).detach()
    errs = []
    for _ in range(reps):
        with torch.no_grad():
            a = m.encode(store.collate(ids, prepare_for="sage", backward=False))
        errs.append(float((a - r).abs().max() / r.abs().max()))
    print(f"sizes={sizes if len(sizes) < 9 else (len(sizes), sizes[0])} rand_bn={rand_bn} layers={layers}: rel err {['%.2e' % e for e in errs]}", flush=True)

run([84] * 9, False)
run([84] * 9, True)
run([84] * 16, False)
run([84] * 16, True)
run([84, 30, 130, 57, 84, 200, 360], False)
run([360] * 3, False)
run([360] * 3, True)
run([200] * 3, True)
run([100] * 3, True)
run([84] * 4, True, layers=2)
run([84] * 1, True, layers=2)
