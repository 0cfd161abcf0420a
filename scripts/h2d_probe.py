"""Host->device copy bandwidth from pinned memory on this box, per CPU affinity / NUMA node (the e2e leg of bench.py is
bound by this number).   python scripts/h2d_probe.py"""
import os, subprocess, time, json
import torch

def bw(nbytes, reps=10):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            d.copy_(h, non_blocking=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps):
            d.copy_(h, non_blocking=True)
        e1.record(s)
    torch.cuda.synchronize()
    return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9

print(subprocess.run("nvidia-smi topo -m; lscpu | grep -i 'numa\\|^CPU(s)\\|model name'; nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv",
                     shell=True, capture_output=True, text=True).stdout)
torch.cuda.init()
all_cpus = sorted(os.sched_getaffinity(0))
print("affinity:", len(all_cpus), "cpus", all_cpus[:4], "...", all_cpus[-4:])
out = {}
for mb in (8, 80, 256):
    out[f"default_{mb}MB"] = bw(mb << 20)
nodes = {}
try:
    for d in sorted(os.listdir("/sys/devices/system/node")):
        if d.startswith("node"):
            cl = open(f"/sys/devices/system/node/{d}/cpulist").read().strip()
            cpus = []
            for part in cl.split(","):
                a, _, b = part.partition("-")
                cpus += list(range(int(a), int(b or a) + 1))
            nodes[d] = [c for c in cpus if c in all_cpus]
except OSError as e:
    print("no /sys node info", e)
for name, cpus in nodes.items():
    if not cpus:
        continue
    os.sched_setaffinity(0, cpus)
    time.sleep(0.05)
    out[f"{name}_80MB"] = bw(80 << 20)      # pinned pages are first-touched by this thread -> on this node
os.sched_setaffinity(0, all_cpus)
print(json.dumps(out, indent=1))
