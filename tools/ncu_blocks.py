"""Summarise an ncu report: per kernel headline metrics, and basic blocks (runs of SASS instructions with equal
execution count) with their share of instructions and stall samples.
usage: python tools/ncu_blocks.py report.ncu-rep kernel-regex [min_pct]"""
import csv
import subprocess
import sys
from collections import Counter

rep, pat = sys.argv[1], sys.argv[2]
minp = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:3]:
    for w in want:
        if w in hdr:
            print(f"{w:70s} {r[hdr.index(w)]}")
    st = sorted(((float(r[hdr.index(h)] or 0), h.split("stalled_")[1].split("_per_issue")[0]) for h in stalls), reverse=True)
    print("stalls/issue:", ", ".join(f"{n}={v:.2f}" for v, n in st[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
first_kernel_done = False
seen, data = set(), []
for r in rows[2:]:
    if len(r) < 9:
        if data:
            break
        continue
    if r[0] in seen:
        break
    seen.add(r[0])
    try:
        data.append((r[1].strip(), int(r[2]), int(r[5]), float(r[8])))
    except ValueError:
        pass
blocks, cur = [], None
for srcl, smp, ex, thr in data:
    op = srcl.split()[0] if not srcl.startswith("@") else srcl.split()[1]
    if cur and cur["ex"] == ex:
        cur["n"] += 1
        cur["smp"] += smp
        cur["ops"].append(op)
    else:
        cur = {"ex": ex, "n": 1, "smp": smp, "thr": thr, "ops": [op]}
        blocks.append(cur)
ti = sum(b["ex"] * b["n"] for b in blocks) or 1
ts = sum(b["smp"] for b in blocks) or 1
print("instructions", ti, "samples", ts)
for b in blocks:
    w = b["ex"] * b["n"]
    if 100 * w / ti > minp or 100 * b["smp"] / ts > minp:
        c = Counter(b["ops"]).most_common(6)
        print(f"exec {b['ex']:8d} x {b['n']:3d} = {100*w/ti:5.1f}% inst {100*b['smp']/ts:5.1f}% smp thr {b['thr']:4.0f} {c}")
