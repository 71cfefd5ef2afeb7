"""Summarise an .ncu-rep (read on the CPU box): per-kernel time, DRAM bytes, throughput, occupancy, top stall reasons."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
def col(r, k):
    return r[hdr.index(k)] if k in hdr else "n/a"
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fma.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio"]
for r in rows[2:]:
    print("==", col(r, "Kernel Name")[:90], "id", col(r, "ID"))
    for k in keys:
        if k in hdr:
            print(f"   {k:70s} {col(r,k)} {rows[1][hdr.index(k)]}")
    stalls = []
    for i, k in enumerate(hdr):
        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
            try:
                stalls.append((float(r[i]), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    print("   stalls:", ", ".join(f"{n}={v:.2f}" for v, n in stalls[:6]))
