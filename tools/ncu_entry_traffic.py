"""ncu raw CSV (``ncu -i X.ncu-rep --page raw --csv``) of ONE bench step -> per-kernel summary lines and the DRAM bytes
per C-ABI call that bench.py reports as ``roofline.traffic`` (profiles/traffic.json).

    python tools/ncu_entry_traffic.py gpurun_out/r02_full_raw.csv profiles/r02_ncu_full_metrics.txt profiles/traffic.json
"""
import csv, json, re, sys

ENTRY = [("k_first_fwd<0", "cgnn_gcn_layer_fwd"), ("k_first_fwd<1", "cgnn_sage_layer_fwd"), ("k_gcn_fwd_ws", "cgnn_gcn_layer_fwd"),
         ("k_gather<0", "cgnn_sage_layer_fwd"), ("k_sage_fwd_gemm", "cgnn_sage_layer_fwd"), ("k_gather<1", "cgnn_gcn_layer_bwd"),
         ("k_gcn_bwd_gemm", "cgnn_gcn_layer_bwd"), ("k_first_bwd<0", "cgnn_gcn_layer_bwd"), ("k_gather<2", "cgnn_sage_layer_bwd"),
         ("k_sage_bwd_gemm", "cgnn_sage_layer_bwd"), ("k_first_bwd<1", "cgnn_sage_layer_bwd"), ("k_collate", "cgnn_collate_csr"),
         ("k_pool_fwd", "cgnn_pool_fwd"), ("k_bn_bwd_sums", "cgnn_bn_bwd_sums"), ("k_head_fwd", "cgnn_head_fwd"), ("k_ce_fwd", "cgnn_ce_fwd")]
CALLS = {"cgnn_gcn_layer_fwd": 6, "cgnn_sage_layer_fwd": 6, "cgnn_gcn_layer_bwd": 3, "cgnn_sage_layer_bwd": 3, "cgnn_collate_csr": 4,
         "cgnn_pool_fwd": 3, "cgnn_bn_bwd_sums": 2, "cgnn_head_fwd": 4, "cgnn_ce_fwd": 4}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}


def main(src, out_txt, out_json):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}

    def val(r, h, want=None):
        v = r[idx[h]].replace(",", "")
        try:
            x = float(v)
        except ValueError:
            return float("nan")
        return x * SCALE.get(units[idx[h]], 1.0)

    stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    lines, per_entry, seen = [], {}, {}
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("cgnn::", "").replace("eng::", "")
        entry = next((e for k, e in ENTRY if k in name), None)
        t = val(r, "gpu__time_duration.sum")
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        stalls = sorted(((h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], val(r, h)) for h in stall_cols), key=lambda x: -x[1])[:4]
        lines.append(f"== {name}  [{entry}]\n   time={t:.1f}us rd={rd / 1e6:.1f}MB wr={wr / 1e6:.1f}MB "
                     f"dram_gbs={(rd + wr) / t / 1e3:.0f} issue%={val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} "
                     f"occ%={val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} regs={val(r, 'launch__registers_per_thread'):.0f} "
                     f"grid={val(r, 'launch__grid_size'):.0f} smem_wavefronts={val(r, 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'):.0f} "
                     f"stalls={[(a, round(b, 1)) for a, b in stalls]}")
        if entry:
            per_entry.setdefault(entry, []).append((name, rd + wr, t))
    out = {"source": f"{out_txt} (ncu --set full --clock-control none, one launch per kernel of one bench step: 4096 x 360-node "
                     "subjects, hidden 64, 3 layers; dram__bytes_read.sum + dram__bytes_write.sum of the entry point's kernels, "
                     "summed over the step and divided by the entry point's calls per step)", "entries": {}}
    for entry, ks in per_entry.items():
        calls = CALLS[entry]
        # the capture window may run one or two launches into the next step: keep a whole number of steps per kernel type
        by = {}
        for name, b, t in ks:
            by.setdefault(name, []).append(b)
        total = 0.0
        detail = []
        for name, bs in by.items():
            per_step = {"cgnn_collate_csr": 4}.get(entry)
            n = per_step if per_step and len(bs) >= per_step else len(bs)
            if entry in ("cgnn_gcn_layer_fwd",) and "first_fwd" in name:
                n = min(len(bs), 2)
            total += sum(bs[:n])
            detail.append(f"{n} x {name} ({sum(bs[:n]) / n / 1e6:.0f} MB)")
        out["entries"][entry] = {"dram_bytes_per_launch": int(total / calls), "calls": " + ".join(detail) + f" over {calls} calls"}
    open(out_txt, "w").write("\n".join(lines) + "\n")
    json.dump(out, open(out_json, "w"), indent=1)
    for e, v in out["entries"].items():
        print(e, v["dram_bytes_per_launch"] / 1e6, "MB/call;", v["calls"])


if __name__ == "__main__":
    main(*sys.argv[1:4])
