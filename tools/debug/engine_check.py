"""GPU check of the warp-specialised forward engine: parity against the previous kernels and per-layer timing.
usage: python tools/debug/engine_check.py [subjects] [regions]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "connectome-gnn-suite_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
from connectome_gnn import _engine
from connectome_gnn._engine import Act
from connectome_gnn.graph import SubjectStore, pack_graphs
from connectome_gnn.synthetic import generate_dataset

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 360
dev = torch.device("cuda", 0)
pool = generate_dataset(num_subjects=min(B, 64), num_regions=N, seed=42)
graphs = (pool * (-(-B // len(pool))))[:B]
store = SubjectStore(pack_graphs(graphs), dev)
batch = store.collate(np.arange(B), prepare_for="gcn")
eng = _engine.engine_for(batch.node_features)
rows = batch.num_nodes
torch.manual_seed(0)
t_in = torch.randn(rows, 64, device=dev)
W = torch.randn(64, 64, device=dev) * 0.2
bias = torch.randn(64, device=dev) * 0.1
scale, shift = torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev) * 0.1
for p_drop in (0.0, 0.3):
    act = Act(scale, shift, True, p_drop, 1234, 1, 0)
    res = {}
    for on in (1, 0):
        eng.lib.cgnn_set_option(1, on)
        z, stats, _ = eng.layer_fwd("gcn", t_in, act, W, bias, batch.csr, batch.ptr, B, want_stats=True)
        torch.cuda.synchronize()
        res[on] = (z.clone(), stats.clone())
    eng.lib.cgnn_set_option(1, 1)
    dz = float((res[1][0] - res[0][0]).abs().max() / res[0][0].abs().max())
    ds = float((res[1][1] - res[0][1]).abs().max() / res[0][1].abs().max())
    print(f"p_drop {p_drop}: z max-norm rel diff {dz:.3e}, stats rel diff {ds:.3e}, count {float(res[1][1][0])} vs {float(res[0][1][0])}", flush=True)
act = Act(scale, shift, True, 0.3, 1234, 1, 0)
for on in (1, 0):
    eng.lib.cgnn_set_option(1, on)
    for want in (True, False):
        for _ in range(3):
            eng.layer_fwd("gcn", t_in, act, W, bias, batch.csr, batch.ptr, B, want_stats=want)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            eng.layer_fwd("gcn", t_in, act, W, bias, batch.csr, batch.ptr, B, want_stats=want)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        gb = (8 * rows * 64 + 4 * (rows + 1) + 8 * 8 * rows + 4 * rows) / 1e9
        print(f"engine={on} stats={want}: {ms*1e3:.1f} us per layer call ({B} x {N}-node), {gb/ms*1e3:.0f} GB/s algorithmic", flush=True)
eng.lib.cgnn_set_option(1, 1)
