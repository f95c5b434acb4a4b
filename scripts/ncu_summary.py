"""Summarise an .ncu-rep (raw + source pages) into text: key metrics, executed opcode mix,
lane efficiency per hot region.  Usage: python scripts/ncu_summary.py report.ncu-rep [out.txt]"""
import csv, subprocess, sys, io
from collections import Counter

rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
for k_i, r in enumerate(rows[2:]):
    print("== kernel launch %d: %s" % (k_i, r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""), file=out)
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print("%-90s %-12s %s" % (k, units[i], r[i]), file=out)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# first kernel only
start = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[start]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[start + 1:]:
    if len(r) < len(hdr) or r[0] in ("Kernel Name", "Address"):
        break
    data.append(r)
tot = sum(int(r[ix["Instructions Executed"]]) for r in data)
thr = sum(int(r[ix["Thread Instructions Executed"]]) for r in data)
print("== source page: SASS lines %d, warp instructions %d, avg active threads %.2f" % (len(data), tot, thr / max(tot, 1)), file=out)
c = Counter()
for r in data:
    op = [o for o in r[ix["Source"]].split() if not o.startswith("@")]
    c[op[0].split(".")[0] if op else "?"] += int(r[ix["Instructions Executed"]])
print("== executed opcode mix (warp instructions)", file=out)
for k, v in c.most_common(28):
    print("  %-10s %12d %5.1f%%" % (k, v, 100.0 * v / tot), file=out)
pop = Counter()
nmax = max(int(x[ix["Instructions Executed"]]) for x in data)
for r in data:
    n = int(r[ix["Instructions Executed"]])
    if n * 50 > nmax:
        pop[(n, r[ix["Avg. Threads Executed"]])] += 1
print("== regions: (executions, avg threads) -> #SASS instructions", file=out)
for k, v in sorted(pop.items(), key=lambda kv: -kv[0][0] * kv[1])[:16]:
    print("  ", k, v, file=out)
