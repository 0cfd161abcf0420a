"""End-to-end run of the reference's per-subject pipeline (Dataload_eeg.py:173-270) on synthetic
data of the dataset's shape: for every subject  raw -> preprocess (GPU) -> 280/120 split -> train
EEGNet_tor -> test accuracy, with the subjects of this rank trained in lock-step.
    python scripts/train_all_subjects.py --subjects 42 --epochs 5            # one GPU
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 scripts/train_all_subjects.py --subjects 42
Subject s goes to rank (s-1) % world (no data-path collective); rank 0 gathers the accuracies.
With --mat-dir DIR the recordings are read from DIR/subjectNN/EEG/subjectNN_eeg.mat (the dataset layout,
Dataload_eeg.py:56-79) through eav_b200.mat_ingest (prefetch thread + copy stream) instead of being synthesised;
--legacy-order selects the paper's band-pass-first preprocessing with labels 0..4 (CNN_EEG_tf.py:180-206)."""
import argparse, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scipy.signal import butter
from eav_b200 import ops
from eav_b200.Dataload_eeg import decimation_taps, epoch_slots
from eav_b200.EAV_datasplit import EAVDataSplit
from eav_b200.sharding import gather_results, subjects_for_rank, train_subjects

ap = argparse.ArgumentParser()
ap.add_argument("--subjects", type=int, default=42)
ap.add_argument("--epochs", type=int, default=5)
ap.add_argument("--lr", type=float, default=1e-5)
ap.add_argument("--separable", action="store_true", help="add a class-dependent component so accuracy can rise above chance")
ap.add_argument("--mat-dir", default=None, help="read subjectNN/EEG/subjectNN_eeg.mat files instead of synthesising")
ap.add_argument("--legacy-order", action="store_true")
a = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    torch.distributed.init_process_group("nccl", device_id=dev)
mine = subjects_for_rank(range(1, a.subjects + 1), rank, world)
S = len(mine)
t0 = time.perf_counter()
data = {}


def split_subject(s, ep, y):
    tr, te = EAVDataSplit(np.zeros((len(y), 1)), y).get_split_indices(h_idx=56)
    data[s] = (ep[torch.from_numpy(tr).to(dev)], torch.from_numpy(y[tr]), ep[torch.from_numpy(te).to(dev)], torch.from_numpy(y[te]))


if a.mat_dir:
    from eav_b200.mat_ingest import prepare_subjects
    t_gen = 0.0
    band = [3, 50] if a.legacy_order else [0.5, 45]
    for s, ep, y in prepare_subjects(a.mat_dir, mine, band=band, device=dev, legacy_order=a.legacy_order):
        split_subject(s, ep.clone(), y if a.legacy_order else (y - 1) // 2)      # harness remap to 0..4
    torch.cuda.synchronize(); t_pre = time.perf_counter() - t0
else:
    # synthetic raw EEG on the device (SURVEY 8d recipe), labels on the host
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    raw = torch.empty(S, 200, 30, 10000, device=dev)
    tt = (torch.arange(2_000_000, device=dev, dtype=torch.float64) / 500.0).reshape(200, 1, 10000)
    wave = (0.5 * torch.sin(2 * np.pi * 50.0 * tt) + 5.0 * torch.sin(2 * np.pi * 0.1 * tt)).float()
    labels = []
    for i, s in enumerate(mine):
        raw[i].normal_(generator=g); raw[i] += wave
        cls = np.random.default_rng(1000 + s).permutation(np.repeat(np.arange(10), 20))
        lab = np.zeros((10, 200)); lab[cls, np.arange(200)] = 1.0
        labels.append(lab)
        if a.separable:      # a 10 Hz rhythm whose topography depends on the class (something for the model to find)
            topo = torch.randn(10, 30, generator=torch.Generator().manual_seed(7)).to(dev)
            osc = torch.sin(2 * np.pi * 10.0 * tt[0, 0].float())
            raw[i] += 0.6 * topo[torch.from_numpy(cls).to(dev)].unsqueeze(-1) * osc
    torch.cuda.synchronize(); t_gen = time.perf_counter() - t0
    plans = [epoch_slots(l) for l in labels]
    order = 1 if a.legacy_order else 0
    eng = ops.PreprocEngine(S, device=dev, order=order)
    sos = butter(5, [3, 50], btype="band", fs=500, output="sos") if a.legacy_order else \
        butter(5, [0.5, 45], btype="bandpass", fs=100, output="sos")
    t1 = time.perf_counter()
    epochs = eng.run(raw, decimation_taps(5), sos, torch.from_numpy(np.stack([p[0] for p in plans])).to(dev), 400)
    torch.cuda.synchronize(); t_pre = time.perf_counter() - t1
    del raw
    for i, s in enumerate(mine):
        split_subject(s, epochs[i], (plans[i][1] - 1) // 2)                       # harness remap to 0..4
t2 = time.perf_counter()
acc, losses = train_subjects(mine, lambda s: data[s], lr=a.lr, batch_size=32, num_epochs=a.epochs, device=dev)
torch.cuda.synchronize(); t_train = time.perf_counter() - t2
all_acc = gather_results(acc)
if rank == 0:
    rec = {"subjects": a.subjects, "world": world, "subjects_this_rank": S, "epochs": a.epochs, "lr": a.lr,
           "source": "mat files" if a.mat_dir else "synthetic", "legacy_order": a.legacy_order,
           "synth_s": t_gen, "preprocess_s": t_pre, "train_s": t_train, "train_s_per_epoch": t_train / a.epochs,
           "train_samples_per_s": S * 280 * a.epochs / t_train,
           "mean_test_acc": float(np.mean(list(all_acc.values()))), "first_epoch_loss": float(np.mean([l[0] for l in losses.values()])),
           "last_epoch_loss": float(np.mean([l[-1] for l in losses.values()])),
           "projected_200_epochs_s": t_train / a.epochs * 200}
    print(json.dumps(rec))
if world > 1:
    torch.distributed.destroy_process_group()
