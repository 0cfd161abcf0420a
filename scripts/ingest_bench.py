"""`.mat` ingest -> device: seconds per subject of
  (a) the reference's way: scipy.io.loadmat, transpose, H2D, then the GPU preprocessing, one subject after another;
  (b) eav_b200.mat_ingest.prepare_subjects: zero-copy MAT v5 reader + pinned float32 staging + H2D on a copy stream,
      overlapped with the GPU preprocessing of the previous subject.
Writes K synthetic dataset-shaped subjects (float64 `seg` (10000, 30, 200), 480 MB each) under --dir first."""
import argparse, json, os, sys, time
import numpy as np, scipy.io, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
from eav_b200 import mat_ingest as MI
from eav_b200.Dataload_eeg import DataLoadEEG

ap = argparse.ArgumentParser()
ap.add_argument("--dir", default="/tmp/eav_mat")
ap.add_argument("--subjects", type=int, default=4)
ap.add_argument("--compress", action="store_true")
a = ap.parse_args()
import eeg_oracle as O
t0 = time.perf_counter()
for s in range(1, a.subjects + 1):
    folder = os.path.join(a.dir, f"subject{s:02d}", "EEG")
    os.makedirs(folder, exist_ok=True)
    raw, label = O.synth_subject(s)
    scipy.io.savemat(os.path.join(folder, f"subject{s:02d}_eeg.mat"), {"seg": np.transpose(raw.astype(np.float64), (2, 1, 0))},
                     do_compression=a.compress)
    scipy.io.savemat(os.path.join(folder, f"subject{s:02d}_eeg_label.mat"), {"label": label})
write_s = time.perf_counter() - t0
subs = list(range(1, a.subjects + 1))
torch.zeros(1, device="cuda")

def reference_way():
    outs = []
    for s in subs:
        D = DataLoadEEG(subject=s, band=[0.5, 45], parent_directory=a.dir)
        mat = scipy.io.loadmat(os.path.join(a.dir, f"subject{s:02d}", "EEG", f"subject{s:02d}_eeg.mat"))
        lab = scipy.io.loadmat(os.path.join(a.dir, f"subject{s:02d}", "EEG", f"subject{s:02d}_eeg_label.mat"))["label"]
        D.set_raw(np.array(mat["seg"]), lab)                 # Dataload_eeg.py:70-82
        x, y = D.prepare_data_device()
        outs.append(x.clone())
    torch.cuda.synchronize()
    return outs

def pipeline():
    outs = []
    for s, x, y in MI.prepare_subjects(a.dir, subs, band=[0.5, 45]):
        outs.append(x.clone())
    torch.cuda.synchronize()
    return outs

res = {"subjects": a.subjects, "compressed": a.compress, "write_s": write_s, "host_cores": len(os.sched_getaffinity(0))}
for name, fn in (("loadmat_sequential", reference_way), ("prefetch_pipeline", pipeline)):
    fn()                                  # warm the page cache and the CUDA context
    t0 = time.perf_counter()
    outs = fn()
    res[name + "_s_per_subject"] = (time.perf_counter() - t0) / a.subjects
    res[name + "_checksum"] = float(sum(float(o.double().abs().sum()) for o in outs))
res["speedup"] = res["loadmat_sequential_s_per_subject"] / res["prefetch_pipeline_s_per_subject"]
print(json.dumps(res))
