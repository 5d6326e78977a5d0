"""Print the hottest SASS instructions (by warp-stall samples) of one kernel from an .ncu-rep."""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
H = rows[h]
si, ii, src = H.index("# Samples"), H.index("Instructions Executed"), H.index("Source")
body = [r for r in rows[h + 1:] if len(r) > si and r[si].isdigit()]
tot = sum(int(r[si]) for r in body)
print(f"kernel {rows[0][1][:80]}  total samples {tot}  instructions {len(body)}")
stall_cols = [i for i, c in enumerate(H) if c.startswith("stall_") and "Not Issued" not in c]
for n, r in sorted(enumerate(body), key=lambda x: -int(x[1][si]))[:top]:
    st = sorted(((H[i][6:], int(r[i] or 0)) for i in stall_cols), key=lambda x: -x[1])[:2]
    print(f"{n:5d} {int(r[si]):6d} {100*int(r[si])/max(tot,1):5.1f}%  exec={r[ii]:>8s}  {r[src].strip()[:70]:70s} {st}")
