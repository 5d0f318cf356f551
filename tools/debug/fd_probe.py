"""Finite-difference probe of the dropout gradients on the GPU: tensor-core vs generic kernels, several seeds and steps."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "connectome-gnn-suite_b200"), os.path.join(ROOT, "tests")]
import torch, helpers, parity
from connectome_gnn import _engine
from connectome_gnn.graph import collate_graphs
from connectome_gnn.train import CrossEntropyLoss
a = helpers.golden("ref_small.npz")
b = collate_graphs(helpers.graphs_from_store(a))
eng = _engine.engine_for(b.node_features)
for kind in ("sage", "gcn"):
    for tc in (1, 0):
        eng.lib.cgnn_set_option(1, tc)
        m = parity.make_model(kind, a, "cuda", dropout=0.3)
        m.train()
        for bn in m.batch_norms: bn.momentum = 0.0
        params = list(m.parameters())
        for seed in (123, 124, 125, 126):
            def loss_at():
                torch.manual_seed(seed)
                return CrossEntropyLoss()(m(b), b.labels)
            m.zero_grad(); l0 = loss_at(); l0.backward()
            norm = float(torch.sqrt(sum((p.grad.double() ** 2).sum() for p in params)))
            direction = [(p.grad / norm).clone() for p in params]
            out = []
            for eps in (1e-3, 2e-4, 5e-5):
                with torch.no_grad():
                    for p, d in zip(params, direction): p.add_(eps * d)
                    lp = float(loss_at())
                    for p, d in zip(params, direction): p.sub_(2 * eps * d)
                    lm = float(loss_at())
                    for p, d in zip(params, direction): p.add_(eps * d)
                out.append((lp - lm) / (2 * eps))
            print(kind, "tc" if tc else "simt", seed, "|g| %.5f" % norm, ["%.5f" % o for o in out], flush=True)
eng.lib.cgnn_set_option(1, 1)
