"""Summarise the per-instruction warp-stall samples of one kernel from `ncu --page source --csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if len(r) != len(hdr) or r[0] == "Address":
        if data:
            break          # first kernel instance only
        continue
    data.append(r)
I = lambda r, k: int(r[idx[k]] or 0)
tot = sum(I(r, "# Samples") for r in data)
reasons = [h for h in hdr if h.startswith("stall_") and "(Not Issued)" in h]
agg = {h: sum(I(r, h) for r in data) for h in reasons}
s = sum(agg.values())
print(f"instructions {len(data)}  samples {tot}  not-issued {s} ({s / max(tot, 1):.3f})")
for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
    print(f"  {h[6:-13]:18s} {v:8d} {v / max(s, 1):.3f}")
byop = {}
for r in data:
    t = r[idx["Source"]].strip().split()
    if not t:
        continue
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    e = byop.setdefault(op, [0, 0, 0])
    e[0] += I(r, "# Samples"); e[1] += I(r, "Instructions Executed"); e[2] += I(r, "Warp Stall Sampling (Not-issued Samples)")
print("by opcode (samples, share, not-issued, executed):")
for op, (a, n, ni) in sorted(byop.items(), key=lambda kv: -kv[1][0])[:10]:
    print(f"  {op:10s} {a:8d} {a / max(tot, 1):.3f} {ni:8d} {n}")
print("top not-issued instructions:")
for r in sorted(data, key=lambda r: -I(r, "Warp Stall Sampling (Not-issued Samples)"))[:10]:
    print("  ", I(r, "Warp Stall Sampling (Not-issued Samples)"), r[idx["Source"]].strip()[:48],
          {h[6:-13]: I(r, h) for h in reasons if I(r, h) > 40})
