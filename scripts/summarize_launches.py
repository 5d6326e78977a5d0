"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean duration, share."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
d = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > vi:
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
        d[r[ki][:72]].append(v)
tot = sum(sum(v) for v in d.values())
print(f"{'kernel':72s} {'n':>5s} {'mean us':>9s} {'share':>7s}")
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:72s} {len(v):5d} {sum(v)/len(v):9.2f} {sum(v)/tot:7.3f}")
print(f"total {tot/1e3:.3f} ms over {sum(len(v) for v in d.values())} launches")
