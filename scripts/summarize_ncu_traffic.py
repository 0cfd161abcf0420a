"""ncu launch list with gpu__time_duration.sum + dram__bytes_{read,write}.sum (CSV, --log-file) -> per-kernel summary CSV
and profiles/traffic.json (DRAM bytes per launch of the kernels bench.py's roofline names).
    python scripts/summarize_ncu_traffic.py gpurun_out/r2_step_launches_raw.csv profiles/r2_step_launches.csv profiles/traffic.json"""
import collections, csv, json, sys

src, dst, tj = sys.argv[1:4]
lines = [l for l in open(src) if not l.startswith("==")]
agg = collections.OrderedDict()
unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for r in csv.DictReader(lines):
    name = r["Kernel Name"].split("(")[0].replace("void ", "")
    v, u = float(r["Metric Value"].replace(",", "")), r["Metric Unit"]
    d = agg.setdefault(name, {"n": 0, "t": 0.0, "rd": 0.0, "wr": 0.0})
    if r["Metric Name"] == "gpu__time_duration.sum":
        d["n"] += 1
        d["t"] += v / 1e3 if u.startswith("n") else v if u.startswith("u") else v * 1e3
    elif r["Metric Name"] == "dram__bytes_read.sum":
        d["rd"] += v * unit.get(u, 1)
    elif r["Metric Name"] == "dram__bytes_write.sum":
        d["wr"] += v * unit.get(u, 1)
ours = {k: d for k, d in agg.items() if "eav::" in k}
tot = sum(d["t"] for d in ours.values())
with open(dst, "w", newline="") as f:
    w = csv.writer(f)                       # kernel names contain commas (template arguments): quoted
    w.writerow(["kernel", "launches", "mean_us", "share_of_eav_time", "dram_read_MB_per_launch", "dram_write_MB_per_launch"])
    for k, d in sorted(ours.items(), key=lambda kv: -kv[1]["t"]):
        w.writerow([k, d["n"], f"{d['t'] / d['n']:.1f}", f"{d['t'] / tot:.4f}", f"{d['rd'] / d['n'] / 1e6:.1f}", f"{d['wr'] / d['n'] / 1e6:.1f}"])
print(open(dst).read())


def per_launch(sub):
    ks = [d for k, d in ours.items() if sub in k]
    return sum((d["rd"] + d["wr"]) / d["n"] for d in ks) if ks else None


step_kernels = [k for k in ours if not any(s in k for s in ("fir_", "sos_", "invert_slots", "needed_trials", "epoch_gather"))]
traffic = {
    "_source": src + " (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch; 42 models x B=32, eval-mode BN step; "
               "preprocessing: 42 subjects, forgetting-filter path)",
    "tconv_fwd": per_launch("tconv_fwd_tc_kernel"),
    "dw_fwd": per_launch("dw_fwd_kernel"),
    "sepconv_fwd": per_launch("sepconv_tc_kernel"),
    "sepconv_bwd_dx": per_launch("sepconv_tc_kernel"),
    "sepconv_bwd_dw": per_launch("sepconv_dw_tc_kernel"),
    "pool1_bwd": per_launch("pool1_bwd_kernel"),
    "tconv_bwd_dw": per_launch("tconv_bwd_fused_tc_kernel"),
    "dw_bwd": 0.0,
    "fir_decimate": per_launch("fir_decimate_kernel"),
    "sos_warm_apply": per_launch("sos_kernel"),
    "preprocess": (per_launch("fir_decimate_kernel") or 0) + (per_launch("sos_kernel") or 0),
    "whole_step": sum((ours[k]["rd"] + ours[k]["wr"]) / ours[k]["n"] * (2 if "sepconv_tc_kernel" in k else 1) for k in step_kernels),
}
json.dump(traffic, open(tj, "w"), indent=1)
print(json.dumps(traffic, indent=1))
