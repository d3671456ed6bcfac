"""One-line-per-metric summary of an `ncu --page raw --csv` export of the rollout kernel (the metrics quoted in profiles/*.md)."""
import csv, sys
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__icc_request_hit_rate.pct",
        "sm__icc_requests.sum", "gcc__cache_requests_type_instruction.sum",
        "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed"]
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    print("==", path)
    for k in WANT:
        if k in d: print(f"  {k:80s} {d[k][0]} {d[k][1]}")
    st = [(h, float(v)) for h, (v, u) in d.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and v]
    print("  stalls per issued instruction:", ", ".join(f"{h[34:-23]} {v:.2f}" for h, v in sorted(st, key=lambda kv: -kv[1])[:8]))
