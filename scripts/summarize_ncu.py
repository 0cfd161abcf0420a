"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into the per-kernel summary
committed under profiles/ (count, mean/total device time, share of all eav:: kernel time)."""
import collections, csv, sys

src, dst = sys.argv[1], sys.argv[2]
lines = [l for l in open(src) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u == "s" else v
    name = row["Kernel Name"].split("(")[0].replace("void ", "")
    agg.setdefault(name, []).append(v)
ours = {k: v for k, v in agg.items() if "eav::" in k}
tot = sum(sum(v) for v in ours.values())
with open(dst, "w", newline="") as f:
    w = csv.writer(f)                       # kernel names contain commas (template arguments): quoted
    w.writerow(["kernel", "launches", "mean_us", "total_us", "share_of_eav_time"])
    for k, v in sorted(ours.items(), key=lambda kv: -sum(kv[1])):
        w.writerow([k, len(v), f"{sum(v)/len(v):.1f}", f"{sum(v):.1f}", f"{sum(v)/tot:.4f}"])
    other = sum(sum(v) for k, v in agg.items() if "eav::" not in k)
    w.writerow(["(non-eav kernels: torch fills/copies/RNG for synthetic data)", sum(len(v) for k, v in agg.items() if "eav::" not in k), "", f"{other:.1f}", ""])
print(open(dst).read())
