"""Debug helper: layer backward, tensor-core path vs generic path, per-output error report."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "connectome-gnn-suite_b200"), os.path.join(ROOT, "tests")]
import torch
from connectome_gnn import _engine
from connectome_gnn._engine import Act, BnBwd
from connectome_gnn.graph import collate_graphs
from connectome_gnn.synthetic import generate_dataset

DEV = torch.device("cuda")
kind = sys.argv[1] if len(sys.argv) > 1 else "gcn"
subjects, regions, d_in, H = [int(v) for v in (sys.argv[2:6] if len(sys.argv) > 5 else (6, 360, 64, 64))]
top = len(sys.argv) > 6 and sys.argv[6] == "top"
eng = _engine.engine_for(torch.zeros(1, device=DEV))
b = collate_graphs(generate_dataset(num_subjects=subjects, num_regions=regions, seed=7))
g = torch.Generator().manual_seed(3)
rows, B = b.num_nodes, b.num_graphs
rn = lambda *s: torch.randn(*s, generator=g)
t_in = rn(rows, d_in).to(DEV)
W = (rn(H, d_in if kind == "gcn" else 2 * d_in) * 0.2).to(DEV)
bias = (rn(H) * 0.1).to(DEV)
act_in = Act((1 + 0.1 * rn(d_in)).to(DEV), (0.1 * rn(d_in)).to(DEV), kind == "gcn", 0.3, seed=79, site=0, row_base=64)
eng.lib.cgnn_set_option(1, 0)
z, _, agg = eng.layer_fwd(kind, t_in, act_in, W, bias, b.csr, b.ptr, B, False)
eng.lib.cgnn_set_option(1, 1)
act_out = Act((1 + 0.1 * rn(H)).to(DEV), (0.1 * rn(H)).to(DEV), kind == "gcn", 0.3, seed=79, site=1, row_base=64)
mean, rstd = (0.1 * rn(H)).to(DEV), (1 + 0.1 * rn(H)).abs().to(DEV)
sums = (rn(2, H) * 0.5).to(DEV)
bn = BnBwd(act_out.scale, mean, rstd, sums, float(rows), True)
du = None if top else rn(rows, H).to(DEV)
demb = rn(B, H).to(DEV) if top else None
pmean, prstd = (0.1 * rn(d_in)).to(DEV), (1 + 0.1 * rn(d_in)).abs().to(DEV)
out = {}
for use_tc in (1, 0):
    eng.lib.cgnn_set_option(1, use_tc)
    out[use_tc] = eng.layer_bwd(kind, du, demb, z, act_out, bn, t_in, act_in, W, b.csr, b.ptr, B, True, pmean, prstd, agg)
    torch.cuda.synchronize()
eng.lib.cgnn_set_option(1, 1)
for name, got, ref in zip(("dW", "dbias", "du_in", "prev_sums"), out[1], out[0]):
    got, ref = got.double().cpu(), ref.double().cpu()
    err = (got - ref).abs().max() / ref.abs().max()
    print(f"{kind} {name:10s} rel err {float(err):.3e}  |got| {float(got.abs().max()):.4g} |ref| {float(ref.abs().max()):.4g} nonfinite {int((~torch.isfinite(got)).sum())}")
    if name == "dW" and err > 1e-4:
        print(" got[0,:8]", got[0, :8].tolist()); print(" ref[0,:8]", ref[0, :8].tolist())
        ratio = got / ref
        print(" ratio stats", float(ratio.median()), float(ratio.min()), float(ratio.max()))
        # does got match a permutation / transposition?
        if got.shape[0] == got.shape[1]:
            print(" vs ref.T", float((got - ref.T).abs().max() / ref.abs().max()))
