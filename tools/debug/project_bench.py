"""Debug helper: time cgnn_project_tf32x3 (P = X W^T, 3xTF32) at the bench's row count.  Run once with
CGNN_PROJECT_TS=0 (operands in shared memory) and once with CGNN_PROJECT_TS=1 (A operand written to tensor memory)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "connectome-gnn-suite_b200")]
import torch
from connectome_gnn import _engine
DEV = torch.device("cuda")
eng = _engine.engine_for(torch.zeros(1, device=DEV))
rows = 4096 * 360
for K, N in ((64, 64), (128, 64)):
    g = torch.Generator().manual_seed(1)
    X = torch.randn(rows, K, generator=g).to(DEV)
    W = (torch.randn(N, K, generator=g) * 0.1).to(DEV)
    P = eng.project_tf32x3(X, W)
    ref = (X[:4096].double() @ W.double().T)
    err = float((P[:4096].double() - ref).abs().max() / ref.abs().max())
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize(); ev[0].record()
    for _ in range(10): eng.project_tf32x3(X, W)
    ev[1].record(); torch.cuda.synchronize()
    us = 100.0 * ev[0].elapsed_time(ev[1])
    gb = rows * (K + N) * 4 / 1e9
    print(f"CGNN_PROJECT_TS={os.environ.get('CGNN_PROJECT_TS', '0')} K={K} N={N}: {us:.1f} us per call, {gb / (us * 1e-6) / 1e3:.2f} TB/s of X + P, rel err {err:.2e}")
