"""Print the headline metrics + top stall instructions of the kernels in an .ncu-rep (read on the CPU box)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_atom.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
units = rows[1]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if pat and pat not in name:
        continue
    print("===", name[:90])
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:85s} {r[i]:>16s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["--kernel-name", "regex:" + pat] if pat else []),
                     capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
if len(rows) > 2:
    hdr = rows[1]
    si, ii, sc, ti = (hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed"),
                      hdr.index("Source"), hdr.index("Avg. Threads Executed"))
    data = []
    for idx, r in enumerate(rows[2:]):
        if len(r) < len(hdr) or r[0] in ("Kernel Name", "Address"):
            break
        data.append((idx, int(r[si] or 0), int(r[ii] or 0), r[sc].strip(), r[ti]))
    tot = sum(d[1] for d in data) or 1
    print(f"--- source: {len(data)} SASS instr, {sum(d[2] for d in data)} warp-instr executed, {tot} stall samples")
    for d in sorted(data, key=lambda x: -x[1])[:14]:
        print(f"  {d[0]:5d} {100 * d[1] / tot:5.1f}%  exec={d[2]:9d} thr={d[4]:>4s}  {d[3][:80]}")
