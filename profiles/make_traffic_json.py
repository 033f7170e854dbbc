"""profiles/r2_ncu_traffic.json from an `ncu --set full --page raw --csv` dump of the fused scan
(tools/r2_profile2.sh): per group count G the mean over the captured launches of DRAM bytes,
shared-memory wavefronts and warp instructions per 32 windows, issue-active and the pipe loads.

    python profiles/make_traffic_json.py profiles/r2_ncu_fused_scan_raw.csv <commit> [bp]
"""
import csv, json, re, sys

src, commit = sys.argv[1], sys.argv[2]
bp = float(sys.argv[3]) if len(sys.argv) > 3 else 1e9
rows = list(csv.reader(open(src)))
h = rows[0]
col = {n: i for i, n in enumerate(h)}


def f(r, name):
    v = r[col[name]].replace(",", "")
    return float(v) if v not in ("", "n/a") else float("nan")


units = rows[1]
by_g = {}
for r in rows[2:]:
    m = re.search(r"fast_scan_kernel<(\d+), (\d+), (\d+)>", r[col["Kernel Name"]])
    if not m:
        continue
    by_g.setdefault(int(m.group(1)), []).append(r)
out = {}
for G, rs in sorted(by_g.items()):
    n = len(rs)
    mean = lambda name: sum(f(r, name) for r in rs) / n
    scale = lambda name, unit_col: 1e9 if units[col[name]] == "Gbyte" else (1e6 if units[col[name]] == "Mbyte" else 1.0)
    dram = sum(f(r, "dram__bytes_read.sum") * scale("dram__bytes_read.sum", 0) +
               f(r, "dram__bytes_write.sum") * scale("dram__bytes_write.sum", 0) for r in rs) / n
    ms = mean("gpu__time_duration.sum") * (1e-3 if units[col["gpu__time_duration.sum"]] == "us" else 1.0)
    e = {
        "dram_bytes_per_launch": dram,
        "kernel_ms_under_ncu": ms,
        "smem_wavefronts_per_32_windows": mean("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") / (bp / 32),
        "warp_instructions_per_32_windows": mean("smsp__inst_executed.sum") / (bp / 32),
        "issue_active_pct": mean("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "pipe_alu_pct": mean("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "pipe_fma_pct": mean("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "pipe_xu_pct": mean("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        "smem_wavefronts_pct_of_peak": mean("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        "registers_per_thread": mean("launch__registers_per_thread"),
        "launches": n,
        "captured_at_commit": commit,
        "workload": "bench.py --motifs 12 --bp %d (fused scan, lean histogram mode, one launch = one motif over the sequence)" % bp,
        "source": "%s (ncu --set full --clock-control none)" % src,
    }
    out["fused_scan_1gbp_G%d" % G] = e
# the aggregate bench.py reads: mean over every captured launch
n_all = sum(e["launches"] for e in out.values())
dram_all = sum(e["dram_bytes_per_launch"] * e["launches"] for e in out.values()) / n_all
out["fused_scan_1gbp"] = {
    "dram_bytes_per_launch": dram_all, "dram_bytes_per_base": dram_all / bp, "captured_at_commit": commit,
    "workload": "mean over the %d captured launches (G = %s)" % (n_all, ", ".join(str(g) for g in sorted(by_g))),
    "source": src,
    "note": "250 MB packed bases + 125 MB ambiguity plane read, 137 MB flag plane + per-tile counts written; "
            "algorithmic bytes 0.25 B/bp = 250 MB"}
json.dump(out, sys.stdout, indent=1)
print()
