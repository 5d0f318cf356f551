"""Debug helper: the wide (H = d_in = 256) GCN layer kernels against a dense fp64 evaluation on the device.

usage: wide_check.py [subjects regions]   (prints per-output max-norm relative errors; exit 1 if any > 1e-5)
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "connectome-gnn-suite_b200"), os.path.join(ROOT, "tests")]
import torch
from connectome_gnn import _engine
from connectome_gnn._engine import Act
from connectome_gnn.graph import collate_graphs
from connectome_gnn.synthetic import generate_dataset

DEV = torch.device("cuda")
subjects, regions = [int(v) for v in (sys.argv[1:3] if len(sys.argv) > 2 else (6, 360))]
H = 256
eng = _engine.engine_for(torch.zeros(1, device=DEV))
b = collate_graphs(generate_dataset(num_subjects=subjects, num_regions=regions, seed=7))
g = torch.Generator().manual_seed(3)
rows, B = b.num_nodes, b.num_graphs
rn = lambda *s: torch.randn(*s, generator=g)
t_in = rn(rows, H).to(DEV)
W = (rn(H, H) * 0.1).to(DEV)
bias = (rn(H) * 0.1).to(DEV)
sc_in, sh_in = (1 + 0.1 * rn(H)).to(DEV), (0.1 * rn(H)).to(DEV)
act_in = Act(sc_in, sh_in, True, 0.0)

# dense A^ (fp64) from the device CSR
c = b.csr
rp, col, wn, dinv = c.in_rowptr.long(), c.in_col.long(), c.in_wn.double(), c.dinv.double()
A = torch.zeros(rows, rows, dtype=torch.float64, device=DEV)
dst = torch.repeat_interleave(torch.arange(rows, device=DEV), rp[1:] - rp[:-1])
A.index_put_((dst, col), wn, accumulate=True)
A += torch.diag(dinv * dinv)

bad = 0
def report(name, got, ref):
    global bad
    got, ref = got.double(), ref.double()
    err = float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
    nonfinite = int((~torch.isfinite(got)).sum())
    flag = "" if (err <= 1e-5 and nonfinite == 0) else "   <-- FAIL"
    bad += bool(flag)
    print(f"{name:12s} rel err {err:.3e}  |got| {float(got.abs().max()):.4g} |ref| {float(ref.abs().max()):.4g} nonfinite {nonfinite}{flag}")
    return err

# ---- forward ----------------------------------------------------------------------------------------------------
z, stats, _ = eng.layer_fwd("gcn", t_in, act_in, W, bias, c, b.ptr, B, True)
torch.cuda.synchronize()
u = torch.relu(t_in.double() * sc_in.double() + sh_in.double())
P = A @ u
z_ref = P @ W.double().T + bias.double()
e = report("z", z, z_ref)
if e > 1e-5:
    # isolate: identity weight -> z = A^ u (gather only)
    eye = torch.eye(H, device=DEV)
    z1, _, _ = eng.layer_fwd("gcn", t_in, act_in, eye, torch.zeros(H, device=DEV), c, b.ptr, B, False)
    report("  gather", z1, P)
    X = eng.project_tf32x3(u.float().contiguous(), W) if hasattr(eng, "project_tf32x3") else None
    err_rows = ((z.double() - z_ref).abs().amax(dim=1) > 1e-4 * z_ref.abs().max()).nonzero().flatten()
    err_cols = ((z.double() - z_ref).abs().amax(dim=0) > 1e-4 * z_ref.abs().max()).nonzero().flatten()
    print("  bad rows", err_rows[:16].tolist(), len(err_rows), " bad cols", err_cols[:16].tolist(), len(err_cols))
report("stats.count", stats[:1], torch.tensor([float(rows)], device=DEV))
report("stats.mean", stats[1:1 + H], z_ref.mean(0))
report("stats.M2", stats[1 + H:], ((z_ref - z_ref.mean(0)) ** 2).sum(0))

# ---- backward (no BatchNorm in between: dz = du * [z > 0]) ------------------------------------------------------------
act_out = Act(None, None, True, 0.0)
du = rn(rows, H).to(DEV)
pmean, prstd = (0.1 * rn(H)).to(DEV), (1 + 0.1 * rn(H)).abs().to(DEV)
dW, db, du_in, prev, _ = eng.layer_bwd("gcn", du, None, z, act_out, None, t_in, act_in, W, c, b.ptr, B, True, pmean, prstd)
torch.cuda.synchronize()
dz = du.double() * (z.double() > 0)
dP = A.T @ dz
report("dbias", db, dz.sum(0))
report("dW", dW, dP.T @ u)
du_ref = dP @ W.double()
report("du_in", du_in, du_ref)
dy = du_ref * ((t_in.double() * sc_in.double() + sh_in.double()) > 0)
xh = (t_in.double() - pmean.double()) * prstd.double()
report("prev_sums", prev, torch.stack([dy.sum(0), (dy * xh).sum(0)]))
sys.exit(1 if bad else 0)
