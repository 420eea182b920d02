"""Key metrics of an ncu report: python tests/ncu_summary.py report.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print("kernel", d.get("Kernel Name"), "grid", d.get("launch__grid_size"), "block", d.get("launch__block_size"))
    keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__warps_active.avg.per_cycle_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct", "sass__inst_executed_local_loads",
            "sass__inst_executed_shared_loads", "sass__inst_executed_global_loads", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
    for k in keys:
        if k in d:
            print(f"  {k} = {d[k]} {units[hdr.index(k)]}")
    st = sorted(((float(v), h) for h, v in d.items() if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and v), reverse=True)
    for v, h in st[:8]:
        print(f"  stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:24s} {v:.2f} cycles/issue")
