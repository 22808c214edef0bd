"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total, average, share."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if r and r[0] == "ID":
        hdr, start = r, i + 1
        break
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg, seq = collections.defaultdict(lambda: [0, 0.0]), collections.defaultdict(list)
for r in rows[start:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(r[ui], v)
    name = r[ki].split("(")[0]
    agg[name][0] += 1
    agg[name][1] += v
    seq[name].append(v)
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':36s} {'n':>6s} {'total us':>11s} {'avg us':>9s} {'share':>6s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    print(f"{k[:36]:36s} {v[0]:6d} {v[1]:11.1f} {v[1] / v[0]:9.2f} {v[1] / tot:6.3f}")
print(f"{'total':36s} {sum(v[0] for v in agg.values()):6d} {tot:11.1f}")
for k in sys.argv[3:]:
    print(k, " ".join(f"{v:.0f}" for v in seq[k][:40]))
