"""Per-kernel totals from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
iu = hdr.index("Metric Unit")
tot, cnt = collections.OrderedDict(), collections.Counter()
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    v = {"ns": v / 1e6, "us": v / 1e3, "ms": v, "nsecond": v / 1e6, "usecond": v / 1e3, "msecond": v, "s": v * 1e3}[r[iu]]
    name = r[ik].split("(")[0]
    tot[name] = tot.get(name, 0.0) + v
    cnt[name] += 1
allms = sum(tot.values())
print(f"{'kernel':60s} {'launches':>8s} {'total ms':>10s} {'ms/launch':>10s} {'share':>7s}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k[:60]:60s} {cnt[k]:8d} {v:10.3f} {v / cnt[k]:10.4f} {v / allms:7.3f}")
print(f"{'sum':60s} {sum(cnt.values()):8d} {allms:10.3f}")
