"""Summarise an .ncu-rep (run here, no GPU needed): key raw metrics + executed-instruction mix + stall samples.
usage: ncu_summary.py report.ncu-rep [kernel-regex] [kernel-index]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else "tile_pass"
kidx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
sel = [r for r in rows[2:] if kre in r[hdr.index("Kernel Name")]]
r = sel[kidx]
print("kernel:", r[hdr.index("Kernel Name")][:90])
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct"]
for k in KEYS:
    if k in hdr:
        print(f"  {k:75s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}")
for i, h in enumerate(hdr):
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
        v = float(r[i])
        if v >= 0.2:
            print(f"  stall {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {v:8.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2 = None
data = []
kc = -1
for row in rows:
    if row and row[0] == "Kernel Name":
        kc += 1
        continue
    if row and row[0] == "Address":
        h2 = row
        continue
    if h2 and len(row) == len(h2) and kc == kidx:
        data.append(row)
if data:
    iS, iI, iP = h2.index("Source"), h2.index("Instructions Executed"), h2.index("# Samples")
    tot = sum(int(x[iI]) for x in data)
    ts = sum(int(x[iP]) for x in data) or 1
    ops, smp = collections.Counter(), collections.Counter()
    for x in data:
        t = x[iS].split()
        op = t[1] if t[0].startswith("@") else t[0]
        ops[op] += int(x[iI])
        smp[op] += int(x[iP])
    print(f"  executed warp instructions: {tot}")
    for op, c in ops.most_common(22):
        print(f"    {op:26s} {100 * c / tot:5.1f}% inst   {100 * smp[op] / ts:5.1f}% samples")
