"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list (per kernel: count, total, share)."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg, seq = collections.OrderedDict(), []
for r in rows[1:]:
    name = r[ki].split("(")[0]
    v = float(r[vi].replace(",", ""))
    seq.append((name, v))
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':34s} {'launches':>8s} {'total ms':>10s} {'avg us':>10s} {'share':>7s}")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k[-34:]:34s} {n:8d} {t / 1e6:10.3f} {t / n / 1e3:10.1f} {100 * t / tot:6.1f}%")
print(f"total {tot / 1e6:.3f} ms over {len(seq)} launches")
if "--seq" in sys.argv:
    for n, v in seq:
        print(f"  {n[-30:]:30s} {v / 1e3:10.1f} us")
