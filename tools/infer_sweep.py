"""BASELINE configs[3]: inference sweep over N unique synthetic subjects (84- or 360-node) sharded across the GPUs of one box.

    python tools/infer_sweep.py --subjects 1000000 --regions 84                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 \
        tools/infer_sweep.py --subjects 1000000 --regions 84

Every rank draws its own shard with the fast generator on its GPU (connectome_gnn.synthetic_fast: all subjects unique),
keeps it resident in a SubjectStore and sweeps it once per model family in batches: collate -> eval-mode forward ->
argmax (kept on the device, summed at the end so that every logit is consumed).  Device time by CUDA events, max over ranks.
Inference has no data-path collective; rank 0 prints one JSON line."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "connectome-gnn-suite_b200"))

import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--subjects", type=int, default=1_000_000)
    ap.add_argument("--regions", type=int, default=84)
    ap.add_argument("--hidden", type=int, default=64)
    ap.add_argument("--batch", type=int, default=0, help="subjects per batch (default: ~1.5 M rows)")
    a = ap.parse_args()
    import torch.distributed as dist
    from connectome_gnn.graph import SubjectStore, shard_bounds
    from connectome_gnn.models import GCNConnectome, GraphSAGEConnectome
    from connectome_gnn.synthetic_fast import generate_packed
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lo, hi = shard_bounds(a.subjects, rank, world)
    n_local = hi - lo
    batch = a.batch or max(1, (4096 * 360) // a.regions)
    t0 = time.time()
    packed = generate_packed(n_local, a.regions, seed=1000 + rank, device=dev)
    torch.cuda.synchronize()
    gen_s = time.time() - t0
    store = SubjectStore(packed, dev)
    arena_gb = sum(packed[k].numel() * packed[k].element_size() for k in ("x", "src", "w")) / 1e9
    del packed
    out = {}
    for kind, cls in (("gcn", GCNConnectome), ("sage", GraphSAGEConnectome)):
        torch.manual_seed(3)
        model = cls(in_channels=5, hidden_dim=a.hidden, num_classes=2, num_layers=3, dropout=0.3).to(dev).eval()
        votes = torch.zeros((), dtype=torch.int64, device=dev)

        def sweep(limit=None):
            nonlocal votes
            done = 0
            for b0 in range(0, n_local, batch):
                ids = np.arange(b0, min(b0 + batch, n_local))
                with torch.no_grad():
                    logits = model(store.collate(ids, prepare_for=kind, backward=False))
                votes += logits.argmax(dim=1).sum()
                done += 1
                if limit and done >= limit:
                    break

        sweep(limit=2)          # warm-up
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sweep()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        out[kind] = {"ms": float(ms), "graphs_per_s": a.subjects / (float(ms) / 1e3), "class1_votes_rank0": int(votes)}
    if rank == 0:
        print(json.dumps({"workload": "BASELINE configs[3]: inference sweep, all subjects unique (synthetic_fast), resident in HBM",
                          "subjects": a.subjects, "regions": a.regions, "hidden": a.hidden, "n_gpus": world, "batch": batch,
                          "generate_s_per_rank": round(gen_s, 2), "arena_gb_per_rank": round(arena_gb, 2), **out}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
