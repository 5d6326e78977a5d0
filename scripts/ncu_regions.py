"""Aggregate warp-stall samples of one kernel by SASS address range (bins), with a few marker instructions."""
import csv, subprocess, sys, re
rep, kern = sys.argv[1], sys.argv[2]
binsz = int(sys.argv[3]) if len(sys.argv) > 3 else 200
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
H = rows[h]
si, ii, src = H.index("# Samples"), H.index("Instructions Executed"), H.index("Source")
body = [r for r in rows[h + 1:] if len(r) > si and r[si].isdigit()]
tot = sum(int(r[si]) for r in body); totx = sum(int(r[ii]) for r in body)
print(f"total samples {tot}, warp-instructions executed {totx}, SASS instructions {len(body)}")
mark = re.compile(r"UTCHMMA|UTMALDG|UTMASTG|LDTM|BAR\.SYNC|LDGSTS|UTCBAR|MUFU|STS\.128|STS\.U16|LDS\.U16|LDS\.128|STG|LDG|SHFL|ATOM|RED")
for b in range(0, len(body), binsz):
    chunk = body[b:b + binsz]
    s = sum(int(r[si]) for r in chunk); x = sum(int(r[ii]) for r in chunk)
    kinds = {}
    for r in chunk:
        m = mark.search(r[src])
        if m: kinds[m.group(0)] = kinds.get(m.group(0), 0) + 1
    print(f"[{b:5d},{b+len(chunk):5d}) samples {s:5d} {100*s/max(tot,1):5.1f}%  exec {x:9d} {100*x/max(totx,1):5.1f}%  {dict(sorted(kinds.items(), key=lambda kv:-kv[1])[:6])}")
