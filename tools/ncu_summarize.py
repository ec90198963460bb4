"""Summarise an `ncu --set full` raw CSV (tools/run_ncu.sh: `ncu -i x.ncu-rep --page raw --csv`) of the K1 kernels into the
text summary kept under profiles/ and the per-launch DRAM traffic file bench.py cites (profiles/ncu_traffic.json).

    python tools/ncu_summarize.py gpurun_out/r2_k1_kernels_raw.csv profiles/r2_k1_kernels_ncu_summary.txt profiles/ncu_traffic.json
"""
import csv
import json
import sys

COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__cluster_size"]


def main(raw, out_txt, out_json):
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {c: [i for i, h in enumerate(hdr) if h == c or h.endswith("." + c)][-1] for c in COLS}
    name_i = hdr.index("Kernel Name")
    lines = ["ncu --set full --clock-control none (tools/run_ncu.sh, target tools/ncu_target.py): K1 at M = 96000, d = 768, r = rg = 96, bf16.",
             "Per-launch times under ncu are cold-cache and serialised.  Columns: " + ", ".join(f"{c} [{units[idx[c]]}]" for c in COLS), ""]
    agg = {}
    for r in data:
        name = r[name_i]
        short = name.split("(")[0].split("::")[-1][:60]
        vals = {c: r[idx[c]] for c in COLS}
        lines.append(f"{short:62s} " + "  ".join(f"{c.split('.')[0].split('__')[-1]}={vals[c]}" for c in COLS))
        key = None
        gated = any(t in name for t in ("_kernel<96, 1", "_kernel<96, true", "_kernel<(int)96, (bool)1"))
        if "k1_fwd_sm100_kernel" in name and gated:
            key = "k1_fwd"
        elif "k1_bwd_sm100_kernel" in name and gated:
            key = "k1_bwd_activation_gradients"
        elif "wgrad_sm100_kernel" in name:
            key = "k1_bwd_weight_gradients"
        if key:
            agg.setdefault(key, []).append((float(vals["gpu__time_duration.sum"]), float(vals["dram__bytes_read.sum"]),
                                            float(vals["dram__bytes_write.sum"]), int(float(vals["launch__grid_size"]))))
    open(out_txt, "w").write("\n".join(lines) + "\n")
    doc = {"capture": f"{out_txt} (round 2, ncu --set full, M = 96000, d = 768, r = rg = 96, bf16)",
           "_comment": "dram__bytes_read.sum + dram__bytes_write.sum per K1 call (the backward tile kernel = main launch + cluster-split "
                       "tail launch, summed); bench.py copies these into k1_micro.*.traffic_MB and tags them as NOT measured by the bench run"}
    # the large-gate iterations come first in the capture: one forward, (main + tail) backward tile launches, one weight-gradient GEMM each
    if "k1_fwd" in agg:
        t, rd, wr, _ = agg["k1_fwd"][0]
        doc["k1_fwd"] = {"read_MB": round(rd, 1), "write_MB": round(wr, 1), "ncu_us": round(t, 1)}
    if "k1_bwd_activation_gradients" in agg:
        ls = agg["k1_bwd_activation_gradients"]
        main_l = max(ls[:2], key=lambda x: x[3])
        tail_l = min(ls[:2], key=lambda x: x[3]) if len(ls) > 1 and ls[0][3] != ls[1][3] else (0.0, 0.0, 0.0, 0)
        doc["k1_bwd_activation_gradients"] = {"read_MB": round(main_l[1] + tail_l[1], 1), "write_MB": round(main_l[2] + tail_l[2], 1),
                                              "ncu_us": round(main_l[0] + tail_l[0], 1)}
    if "k1_bwd_weight_gradients" in agg:
        t, rd, wr, _ = agg["k1_bwd_weight_gradients"][0]
        doc["k1_bwd_weight_gradients"] = {"read_MB": round(rd, 1), "write_MB": round(wr, 1), "ncu_us": round(t, 1)}
    json.dump(doc, open(out_json, "w"), indent=1)
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:4])
