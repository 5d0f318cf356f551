"""GCN training step gradients on one GPU against a float64 autograd restatement, for a few batch sizes (dropout 0)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "connectome-gnn-suite_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from connectome_gnn.graph import SubjectStore, pack_graphs
from connectome_gnn.models import GCNConnectome
from connectome_gnn.synthetic import generate_dataset
from connectome_gnn.train import CrossEntropyLoss

pool = generate_dataset(num_subjects=256, num_regions=360, k=8, beta=0.15, trait_idx=0, seed=42)
for n_all in (96, 192, 296, 297, 384):
    graphs = (pool * (n_all // len(pool) + 1))[:n_all]
    store = SubjectStore(pack_graphs(graphs), "cuda")
    torch.manual_seed(77)
    model = GCNConnectome(in_channels=5, hidden_dim=64, num_classes=2, num_layers=3, dropout=0.0).to("cuda").train()
    runs = []
    for rep in range(2):
        model.zero_grad()
        for bn in model.batch_norms:
            bn.reset_running_stats()
        b = store.collate(np.arange(n_all), prepare_for="gcn")
        loss = CrossEntropyLoss()(model(b), b.labels)
        loss.backward()
        runs.append({n: p.grad.detach().clone() for n, p in model.named_parameters()})
    same = all(torch.equal(runs[0][n], runs[1][n]) for n in runs[0])
    # float64 restatement with autograd
    full = store.collate(np.arange(n_all))
    x = full.node_features.double(); ei = full.edge_index; w = full.edge_weight.double(); bt = full.batch
    N = x.shape[0]
    params = {n: p.detach().double().requires_grad_(True) for n, p in model.named_parameters()}
    loops = torch.arange(N, device="cuda")
    src = torch.cat([ei[0], loops]); dst = torch.cat([ei[1], loops]); ww = torch.cat([w, torch.ones(N, dtype=torch.float64, device="cuda")])
    deg = torch.zeros(N, dtype=torch.float64, device="cuda").index_add_(0, src, ww)
    dinv = (deg + 1e-8).pow(-0.5)
    norm = dinv[src] * ww * dinv[dst]
    h = x
    for l in range(3):
        W, bias = params[f"convs.{l}.linear.weight"], params[f"convs.{l}.bias"]
        hw = h @ W.T
        z = torch.zeros(N, W.shape[0], dtype=torch.float64, device="cuda").index_add_(0, dst, hw[src] * norm[:, None]) + bias
        mean = z.mean(0); var = z.var(0, unbiased=False)
        y = (z - mean) / torch.sqrt(var + 1e-5) * params[f"batch_norms.{l}.weight"] + params[f"batch_norms.{l}.bias"]
        h = torch.relu(y)
    cnt = torch.zeros(n_all, dtype=torch.float64, device="cuda").index_add_(0, bt, torch.ones(N, dtype=torch.float64, device="cuda"))
    emb = torch.zeros(n_all, 64, dtype=torch.float64, device="cuda").index_add_(0, bt, h) / (cnt[:, None] + 1e-8)
    hid = torch.relu(emb @ params["classifier.0.weight"].T + params["classifier.0.bias"])
    logits = hid @ params["classifier.3.weight"].T + params["classifier.3.bias"]
    l64 = torch.nn.functional.cross_entropy(logits, full.labels)
    l64.backward()
    gmax = max(float(p.grad.abs().max()) for p in params.values())
    per = {n: float((runs[0][n].double() - params[n].grad).abs().max()) / gmax for n in params}
    worst = max(per, key=per.get)
    print(f"B={n_all}: deterministic={same} loss {float(loss):.7f} vs f64 {float(l64):.7f}; worst gradient vs f64: {worst} {per[worst]:.2e}; conv0.W {per['convs.0.linear.weight']:.2e}", flush=True)
