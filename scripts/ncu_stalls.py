"""Stall-reason totals and the hottest SASS lines of one kernel from an .ncu-rep source page.
usage: python scripts/ncu_stalls.py <rep> <kernel-regex> [top-n]"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
b = blocks[0]
h = b["rows"][0]
si, src = h.index("# Samples"), h.index("Source")
stall = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
data = [r for r in b["rows"][1:] if len(r) > si]
tot = sum(int(r[si] or 0) for r in data)
print(b["name"], "| samples", tot, "| SASS instructions", len(data))
sums = {c: sum(int(r[i] or 0) for r in data) for i, c in stall}
for c, v in sorted(sums.items(), key=lambda kv: -kv[1]):
    if v: print(f"  {c:26s} {v:7d} {v / max(tot, 1):6.3f}")
print("hottest lines:")
for k, r in sorted(enumerate(data), key=lambda kr: -int(kr[1][si] or 0))[:topn]:
    n = int(r[si] or 0)
    why = max(stall, key=lambda ic: int(r[ic[0]] or 0))[1]
    print(f"  #{k:5d} {n:6d} {n / max(tot, 1):6.3f} {why:22s} {r[src].strip()[:90]}")
