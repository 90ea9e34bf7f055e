"""Summarise one kernel of an .ncu-rep (ncu --set full --import-source on): duration, pipe utilisation, warp-stall samples by reason
and by SASS opcode.  usage: python tools/ncu_stalls.py report.ncu-rep [> summary.md]   (runs `ncu -i` here, no GPU needed)"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = dict(zip(hdr, vals))
u = dict(zip(hdr, units))
print(f"# {m.get('Kernel Name', '?')}\n")
print(f"source: `{rep}` (ncu --set full --clock-control none, one launch; cold caches, serialised)\n")
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]
print("| metric | value | unit |\n|---|---:|---|")
for k in keys:
    if k in m and m[k] != "":
        print(f"| `{k}` | {m[k]} | {u.get(k, '')} |")
srows = list(csv.reader(io.StringIO(src)))
sh = srows[1]
ix = {h: i for i, h in enumerate(sh)}
data = srows[2:]
reasons = [h for h in sh if h.startswith("stall_") and "(Not Issued)" not in h]
tot = Counter()
for r in data:
    for h in reasons:
        try:
            tot[h] += int(r[ix[h]] or 0)
        except ValueError:
            pass
total = sum(tot.values()) or 1
print("\n## warp-stall samples by reason\n\n| reason | samples | share |\n|---|---:|---:|")
for h, n in tot.most_common():
    if n:
        print(f"| {h} | {n} | {100 * n / total:.1f}% |")
ops = Counter()
for r in data:
    s = r[ix["Source"]].split()
    op = s[0] if s and not s[0].startswith("@") else (s[1] if len(s) > 1 else "?")
    ops[op.split(".")[0]] += int(r[ix["# Samples"]] or 0)
ts = sum(ops.values()) or 1
print("\n## samples by SASS opcode (top 12)\n\n| opcode | samples | share |\n|---|---:|---:|")
for op, n in ops.most_common(12):
    print(f"| {op} | {n} | {100 * n / ts:.1f}% |")
