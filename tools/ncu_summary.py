"""Summarise an .ncu-rep (raw + source pages) into a small text file for profiles/.
usage: python tools/ncu_summary.py gpurun_out/X.ncu-rep profiles/X_summary.txt"""
import collections
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw[raw.index('"ID"'):])))
hdr, units, vals = rows[0], rows[1], rows[2]
want = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
]  # fmt: skip
lines = ["ncu summary of %s (ncu --set full --clock-control none; per launch)" % rep, ""]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        lines.append("%-85s %s %s" % (w, vals[i], units[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src[src.index('"Address"'):])))
sh, data = srows[0], []
for r in srows[1:]:  # a report with several launches: the first one only (the raw-page row summarised above)
    if len(r) < len(sh):
        break
    data.append(r)
ia, it, isrc, ismp = sh.index("Instructions Executed"), sh.index("Thread Instructions Executed"), sh.index("Source"), sh.index("# Samples")
cnt = collections.Counter()
tot = thr = 0
for r in data:
    n = int(r[ia])
    if not n:
        continue
    toks = r[isrc].split()
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
    cnt[op] += n
    tot += n
    thr += int(r[it])
lines += ["", "dynamic SASS mix (warp-level instructions executed, whole launch): total %d, avg active threads %.2f" % (tot, thr / tot)]
for op, n in cnt.most_common(24):
    lines.append("  %-10s %12d  %5.1f%%" % (op, n, 100.0 * n / tot))
fp64 = sum(cnt[o] for o in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
lines.append("  FP64 arithmetic share of issued instructions: %.1f%%" % (100.0 * fp64 / tot))
stall_cols = [c for c in sh if c.startswith("stall_") and "Not Issued" not in c]
agg = collections.Counter()
for r in data:
    for c in stall_cols:
        v = r[sh.index(c)]
        if v and v != "0":
            agg[c] += int(v)
ts = sum(agg.values())
lines += ["", "warp-state samples (all): " + ", ".join("%s %.1f%%" % (k, 100.0 * v / ts) for k, v in agg.most_common(8))]
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
