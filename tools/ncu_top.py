"""Top SASS instructions of a kernel by stall samples, with neighbours: python tools/ncu_top.py rep kernel-regex [n] [ctx]"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 15
ctx = int(sys.argv[4]) if len(sys.argv) > 4 else 3
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
iS, iN, iE, iT = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[2:] if len(r) == len(hdr) and r[0].startswith("0x")]
seen = set(); uniq = []
for r in data:
    if r[0] in seen: break
    seen.add(r[0]); uniq.append(r)
tot = sum(int(r[iN]) for r in uniq)
order = sorted(range(len(uniq)), key=lambda i: -int(uniq[i][iN]))[:n]
print("total samples", tot, "instructions", len(uniq))
for i in sorted(order):
    print("----")
    for j in range(max(0, i - ctx), min(len(uniq), i + ctx + 1)):
        r = uniq[j]
        st = sorted(((int(r[c] or 0), h[6:]) for c, h in stall_cols), reverse=True)[:3]
        mark = ">>" if j == i else "  "
        print(f"{mark} {j:5d} smp {int(r[iN]):5d} ({100*int(r[iN])/tot:4.1f}%) exec {r[iE]:>8s} thr {r[iT]:>4s}  {r[iS].strip():60s} {st if j==i else ''}")
