#!/usr/bin/env python
"""bench.py -- EEGNet train samples/sec (fwd+bwd+step) and preprocessing GB/s on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[2], the one the metric is quoted on): the reference's 42
per-subject EEGNet_tor models (Chans=30, Samples=500, kernLength=300, F1=8, D=8, F2=64, 5
classes, Adam lr=1e-5, batch 32) trained in lock-step on synthetic EEG of the dataset's shape
(200 x 10000 x 30 per subject).  One "step" = one batch of 32 for each of the 42 models =
1344 samples through forward + loss + backward + Adam.  Raw synthetic EEG is generated on the
device, preprocessed by the CUDA FIR/SOS/epoch pipeline (timed separately: `preprocess`),
split 280/120 with the reference's index logic, and the training set stays resident in HBM.
With N > 1 every rank holds its own 42 subjects (weak scaling, no data-path collective:
subjects are independent); `value` is the aggregate over ranks, time = max over ranks.

`--impl reference` times the CPU restatement of the reference (oracle/, kind "port": the
reference is Python and /root/reference does not exist on the GPU box) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "EEGNet train samples/sec (fwd+bwd+step)"
N_SUBJECTS, BATCH, N_TRAIN, N_TEST = 42, 32, 280, 120
FLOP_PER_SAMPLE = 198.9e6          # fwd 90.3 + bwd 108.6 MFLOP (conv/dense MACs only)
TCONV_FLOP_PER_SAMPLE = 72.0e6     # 2*K1*F1*Chans*Samples = 2*300*8*30*500, fwd; identical for dW1
PREPROC_BYTES_PER_SUBJECT = 264e6  # 240 MB raw f32 read once + 24 MB epochs written once


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--subjects", type=int, default=N_SUBJECTS, help="subject models per GPU")
    ap.add_argument("--bn-mode", default="train", choices=["train", "eval"],
                    help="BN mode of the headline step (the other mode is reported beside it)")
    ap.add_argument("--no-preproc", action="store_true", help="skip the preprocessing leg (epochs drawn N(0,1))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stages", action="store_true", help="skip the per-kernel stage timing")
    return ap.parse_args()


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region.  In-process NVML polling every 10 ms
    (nvidia_ml_py); `nvidia-smi -lms` is the fallback (its start-up can miss short multi-GPU runs)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.nvml, self.samples, self.mask, self.stop_flag = None, [], 0, False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        while not self.stop_flag:
            try:
                self.samples.append(float(self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM)))
                self.mask |= int(self.get_reasons(self.handle))
            except Exception:
                pass
            time.sleep(0.01)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.t.join(timeout=1)
            return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(k for k, b in self.BITS.items() if self.mask & b), "samples": len(self.samples),
                    "how": "NVML polled every 10 ms during the timed region"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            f = [t.strip() for t in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU baseline (oracle port)
def cpu_train_baseline(steps, warmup, bn_train=True):
    """The CPU restatement of the reference's Trainer step (oracle/eegnet_oracle.py) on the host
    cores: B=32 steps of ONE subject model (1/42 of a GPU step)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import eegnet_oracle as EO
    from eav_b200.CNN_torch.EEGNet_tor import EEGNet_tor
    try:        # every host core this process may use (torchrun pins OMP_NUM_THREADS=1 otherwise)
        torch.set_num_threads(len(os.sched_getaffinity(0)))
    except Exception:
        pass
    torch.manual_seed(1)
    sd = EEGNet_tor(5).state_dict()
    params, buffers = EO.split_state(sd, "tor")
    opt = EO.Adam(params, lr=1e-5)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(BATCH, 1, 30, 500, generator=g)
    y = torch.randint(0, 5, (BATCH,), generator=g)
    for _ in range(warmup):
        EO.train_step("tor", params, buffers, opt, x, y, bn_train)
    t0 = time.perf_counter()
    for _ in range(steps):
        EO.train_step("tor", params, buffers, opt, x, y, bn_train)
    dt = time.perf_counter() - t0
    return BATCH * steps / dt, dt / steps * 1e3, torch.get_num_threads()


def cpu_preproc_baseline():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import eeg_oracle as O
    raw, label = O.synth_subject(1)
    O.build_c()
    t0 = time.perf_counter()
    O.prepare_data(raw, label, [0.5, 45])
    dt = time.perf_counter() - t0
    return PREPROC_BYTES_PER_SUBJECT / dt / 1e9, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    budget_steps = min(steps, 60)
    v, ms, threads = cpu_train_baseline(budget_steps, min(warmup, 3), bn_train=args.bn_mode == "train")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "EEGNet_tor per-subject training, Chans=30 Samples=500 kernLength=300 F1=8 D=8 F2=64, "
                                   "5 classes, Adam lr=1e-5, batch 32", "bn_mode": args.bn_mode,
                       "note": "each timed step = one B=32 train step of ONE subject model on the host cores"},
            "cpu_baseline": {"value": v, "unit": "samples/s", "cores": threads, "kind": "port",
                             "sample": f"{budget_steps} B=32 train steps (fwd+bwd+Adam, {args.bn_mode}-mode BN) of one subject model; "
                                       "oracle/eegnet_oracle.py restatement of CNN_torch/EEGNet_tor.py on torch-CPU"},
            "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


# ----------------------------------------------------------------------------- B200 arm
def synth_raw_device(S, seed, device):
    """Synthetic raw EEG on the device, dataset shape [S][200][30][10000] f32 (SURVEY 8d recipe:
    N(0,1) + 0.5 sin(2 pi 50 t) + 5 sin(2 pi 0.1 t), t continuous across trials) + one-hot labels."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    raw = torch.empty(S, 200, 30, 10000, dtype=torch.float32, device=device)
    t = (torch.arange(200 * 10000, device=device, dtype=torch.float64) / 500.0).reshape(200, 1, 10000)
    wave = (0.5 * torch.sin(2 * np.pi * 50.0 * t) + 5.0 * torch.sin(2 * np.pi * 0.1 * t)).float()
    for s in range(S):
        raw[s].normal_(generator=g)
        raw[s] += wave
    labels = []
    for s in range(S):
        rng = np.random.default_rng(1000 + seed * 1000 + s)
        cls = rng.permutation(np.repeat(np.arange(10), 20))
        lab = np.zeros((10, 200))
        lab[cls, np.arange(200)] = 1.0
        labels.append(lab)
    return raw, labels


_REAL_STDOUT = None


def _claim_stdout():
    """Route everything libraries print on fd 1 (e.g. NCCL's version banner) to stderr, so that
    the single JSON line is the only thing this process ever writes to stdout."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def _emit(line):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def main():
    args = parse()
    _claim_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from eav_b200 import _lib, ops
    from eav_b200.Dataload_eeg import decimation_taps, epoch_slots
    from eav_b200.EAV_datasplit import EAVDataSplit
    from eav_b200.CNN_torch.EEGNet_tor import EEGNet_tor
    from eav_b200.trainer_core import SubjectBatchTrainer
    from scipy.signal import butter

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.require_device()
    lib = _lib.load()
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    M, B, K, W = args.subjects, BATCH, args.steps, max(3, args.warmup)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------------------------------------------------------- data: raw -> preprocess -> split
    pre = None
    if not args.no_preproc:
        raw, labels = synth_raw_device(M, rank, dev)
        taps = decimation_taps(5)
        sos = butter(5, [0.5, 45], btype="bandpass", fs=100, output="sos")
        plans = [epoch_slots(l) for l in labels]
        slot = torch.from_numpy(np.stack([p[0] for p in plans])).to(dev)
        eng = ops.PreprocEngine(M, device=dev)
        epochs = torch.empty(M, 400, 30, 500, dtype=torch.float32, device=dev)
        l0 = lib.eav_launch_count()
        eng.run(raw, taps, sos, slot, 400, epochs=epochs)       # warm-up
        pre_launches = lib.eav_launch_count() - l0
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        barrier()
        e0.record()
        for _ in range(reps):
            eng.run(raw, taps, sos, slot, 400, epochs=epochs)
        e1.record()
        torch.cuda.synchronize()
        pre_ms = max_over_ranks(e0.elapsed_time(e1) / reps)
        hbm_peak = None
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
            peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
        except Exception:
            hbm_peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        gbs = PREPROC_BYTES_PER_SUBJECT * M / (pre_ms * 1e-3) / 1e9
        pre = {"metric": "preprocess HBM GB/s (filter/decimate/epoch)", "value": gbs * world, "unit": "GB/s",
               "ms": pre_ms, "subjects_per_gpu": M, "launches": int(pre_launches),
               "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                            "traffic": None, "peak_source": peak_src,
                            "algorithmic_bytes": PREPROC_BYTES_PER_SUBJECT * M}}
        # the legacy order (band-pass at 500 Hz over the raw recording, then decimate: CNN_EEG_tf.py:64-75,180-189) on
        # a bounded number of subjects (it needs one more raw-sized buffer); reported beside the shipped order
        try:
            from scipy.signal import butter as _butter
            ML = min(M, 8)
            eng2 = ops.PreprocEngine(ML, device=dev, order=1)
            sos500 = _butter(5, [3, 50], btype="band", fs=500, output="sos")
            ep2 = torch.empty(ML, 400, 30, 500, dtype=torch.float32, device=dev)
            eng2.run(raw[:ML], taps, sos500, slot[:ML].contiguous(), 400, epochs=ep2)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                eng2.run(raw[:ML], taps, sos500, slot[:ML].contiguous(), 400, epochs=ep2)
            e1.record()
            torch.cuda.synchronize()
            lms = e0.elapsed_time(e1) / reps
            # raw read twice (state pass, apply pass) + filtered written and read + decimated written and read + epochs
            lbytes = ML * (4 * 240e6 + 2 * 48e6 + 24e6)
            pre["bandpass_first_order"] = {"ms": lms, "subjects": ML, "ms_per_subject": lms / ML,
                                           "algorithmic_bytes": lbytes, "achieved_gbs": lbytes / (lms * 1e-3) / 1e9,
                                           "frac_of_hbm_peak": lbytes / (lms * 1e-3) / 1e9 / hbm_peak}
            del eng2, ep2
        except Exception as e:  # noqa: BLE001
            pre["bandpass_first_order"] = {"error": repr(e)}
        # labels {1,3,5,7,9} -> 0..4 (harness remap) and the reference's 280/120 split
        tr_rows, tr_y = [], []
        for s in range(M):
            y = (plans[s][1] - 1) // 2
            tri, _ = EAVDataSplit(np.zeros((400, 1)), y).get_split_indices(h_idx=56)
            tr_rows.append(torch.from_numpy(tri + 400 * s))
            tr_y.append(torch.from_numpy(y[tri]))
        rows = torch.cat(tr_rows).to(dev)
        x_train = epochs.reshape(M * 400, 30, 500).index_select(0, rows).contiguous()
        y_train = torch.cat(tr_y).to(dev)
        del raw, epochs, eng
        torch.cuda.empty_cache()
    else:
        g = torch.Generator(device=dev).manual_seed(rank)
        x_train = torch.randn(M * N_TRAIN, 30, 500, generator=g, device=dev)
        y_train = torch.randint(0, 5, (M * N_TRAIN,), generator=g, device=dev)

    # ---------------------------------------------------------------- models
    dims = ops.EegnetDims(5)
    tr = SubjectBatchTrainer(dims, M, x_train, y_train, lr=1e-5, max_batch=B, seed=1234 + rank)
    sds = []
    for s in range(M):
        torch.manual_seed(1 + s + 1000 * rank)
        sds.append(EEGNet_tor(5).state_dict())
    tr.load_state_dicts(sds, EEGNet_tor._BN_NAMES)

    # index schedule: per model a fresh permutation per epoch, 8 full batches of 32 per epoch
    n_sched = W + K + 8
    gen = torch.Generator().manual_seed(7 + rank)
    sched = torch.empty(n_sched, M * B, dtype=torch.int32)
    perms = None
    for i in range(n_sched):
        if i % 8 == 0:
            perms = torch.stack([torch.randperm(N_TRAIN, generator=gen) for _ in range(M)])
        b = i % 8
        sched[i] = (perms[:, b * B:(b + 1) * B] + torch.arange(M).unsqueeze(1) * N_TRAIN).reshape(-1).int()
    sched = sched.to(dev)

    def timed_steps(bn_train, n_warm, n_steps):
        for i in range(n_warm):
            tr.train_step(sched[i], bn_train=bn_train)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_steps):
            tr.train_step(sched[n_warm + i], bn_train=bn_train)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1) / n_steps)

    head_train = args.bn_mode == "train"
    # the other BN mode first (short), then the headline run with the clock sampler on
    other_ms = timed_steps(not head_train, 3, min(K, 10))
    l0 = lib.eav_launch_count()
    tr.use_graph = False
    tr.program(B, head_train, "train").graph = None
    tr.train_step(sched[0], bn_train=head_train)
    torch.cuda.synchronize()
    launches_per_step = int(lib.eav_launch_count() - l0)
    tr.use_graph = True
    clocks = ClockSampler(local)
    clocks.start()
    ms = timed_steps(head_train, W, K)
    clk = clocks.stop()
    value = world * M * B / (ms * 1e-3)
    loss_now = tr.program(B, head_train, "train").loss.float().mean().item()

    # ---------------------------------------------------------------- end-to-end: host buffers in, loss out
    p = tr.host_step_program(B, bn_train=head_train)
    n_ring = 4
    host_x = [torch.empty(M * B, 30, 500, dtype=torch.float32).pin_memory() for _ in range(n_ring)]
    host_y = [torch.empty(M * B, dtype=torch.int64).pin_memory() for _ in range(n_ring)]
    for r in range(n_ring):
        host_x[r].copy_(x_train.index_select(0, sched[r].long()).cpu())
        host_y[r].copy_(y_train.index_select(0, sched[r].long()).cpu())
    host_loss = torch.empty(M, dtype=torch.float32).pin_memory()
    stage_x = [p.x_src, torch.empty_like(p.x_src)]
    stage_y = [p.y_src, torch.empty_like(p.y_src)]
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream()

    def e2e_steps(n):
        ev_copied = [torch.cuda.Event() for _ in range(2)]
        ev_used = [torch.cuda.Event() for _ in range(2)]
        with torch.cuda.stream(copy_stream):
            stage_x[0].copy_(host_x[0], non_blocking=True); stage_y[0].copy_(host_y[0], non_blocking=True)
            ev_copied[0].record(copy_stream)
        for i in range(n):
            cur, nxt = i & 1, (i + 1) & 1
            if i + 1 < n:                                   # H2D of step i+1 overlaps the kernels of step i
                with torch.cuda.stream(copy_stream):
                    if i >= 1:
                        copy_stream.wait_event(ev_used[nxt])
                    stage_x[nxt].copy_(host_x[(i + 1) % n_ring], non_blocking=True)
                    stage_y[nxt].copy_(host_y[(i + 1) % n_ring], non_blocking=True)
                    ev_copied[nxt].record(copy_stream)
            main_stream.wait_event(ev_copied[cur])
            p.x_src, p.y_src = stage_x[cur], stage_y[cur]
            p.enqueue()
            ev_used[cur].record(main_stream)
            host_loss.copy_(p.loss, non_blocking=True)      # D2H of the step's result
        torch.cuda.synchronize()

    e2e_steps(3)
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_steps(K)
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / K)
    e2e_value = world * M * B / (e2e_ms * 1e-3)
    h2d = M * B * (30 * 500 * 4 + 8)

    # ---------------------------------------------------------------- per-kernel timing (roofline of the dominant kernel)
    stages, roof = None, None
    if not args.no_stages:
        import ctypes
        prog = tr.program(B, head_train, "train")
        cfg = prog.cfg()
        prog.graph = None
        prog.enqueue()                                      # leaves a complete forward/backward in the workspace
        torch.cuda.synchronize()
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        n_st = lib.eav_eegnet_stage_count()
        stages = {}
        reps = 5
        for s_id in range(n_st):
            name = lib.eav_eegnet_stage_name(s_id).decode()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

            def go():
                _lib.check(lib.eav_eegnet_run_stage(ctypes.byref(cfg), s_id, ops._ptr(tr.x), ops._ptr(prog.idx),
                                                    ops._ptr(tr.params), ops._ptr(tr.bn_state), None, None,
                                                    ops._ptr(prog.out), ops._ptr(prog.dout), ops._ptr(tr.grads),
                                                    ops._ptr(tr.workspace), tr.ws_bytes, st), "run_stage")
            go()
            torch.cuda.synchronize()
            a.record()
            for _ in range(reps):
                go()
            b.record()
            torch.cuda.synchronize()
            stages[name] = a.elapsed_time(b) / reps
        tot = sum(stages.values())
        N = M * B
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 7700.0)
        # dense tf32 runs at half the bf16 rate on tcgen05 (kind::tf32 consumes K=8 per 32 B, kind::f16 K=16)
        tf32_peak = peaks.get("bf16_tflops", 2250.0) / 2
        peak_src_t = ("MEASURED_PEAKS.json bf16_tflops (burst) / 2: dense tf32 peak of tcgen05" if peaks else
                      "fallback: nominal 2250 TF/s bf16 / 2")
        fp32_peak = ops.measure_fp32_peak()
        fp32_ffma2 = ops.measure_fp32_peak(3)
        off = ("ffma", "0")
        tc_all = os.environ.get("EAV_TC", "") not in off
        tc_on = tc_all and os.environ.get("EAV_TCONV", "") not in off            # temporal conv on tcgen05
        sc_on = tc_all and os.environ.get("EAV_SEPCONV", "") not in off          # block-2 conv on tcgen05
        conv2 = 2.0 * 64 * 64 * 16 * 125           # block-2 (1,16) conv, flops per sample
        # algorithmic work of each hot kernel, per launch
        model = {
            "tconv_fwd": ("tensor" if tc_on else "fp32", TCONV_FLOP_PER_SAMPLE * N, "flop"),
            "tconv_bwd_dw": ("tensor" if tc_on else "fp32", TCONV_FLOP_PER_SAMPLE * N, "flop"),
            "sepconv_fwd": ("tensor" if sc_on else "fp32", conv2 * N, "flop"),
            "sepconv_bwd_dx": ("tensor" if sc_on else "fp32", conv2 * N, "flop"),
            "sepconv_bwd_dw": ("tensor" if sc_on else "fp32", conv2 * N, "flop"),
            # y1 + dz1 (8x30x500 each) + y2 + dz2 (64x500 each), fp32
            "dw_bwd": ("hbm", 4.0 * N * (2 * 8 * 30 * 500 + 2 * 64 * 500), "byte"),
            # y1 read + y2 write
            "dw_fwd": ("hbm", 4.0 * N * (8 * 30 * 500 + 64 * 500), "byte"),
        }

        def kernel_roofline(name):
            bound, work, kind = model[name]
            t = stages[name] * 1e-3
            if kind == "flop":
                ach, unit = work / t / 1e12, "TFLOP/s"
                pk = tf32_peak if bound == "tensor" else fp32_peak
            else:
                ach, unit = work / t / 1e9, "GB/s"
                pk = hbm_peak
            return {"kernel": name, "bound": bound, "achieved": ach, "peak": pk, "unit": unit, "frac": ach / pk,
                    "ms_per_launch": stages[name], "share_of_step": stages[name] / tot,
                    ("algorithmic_flops_per_launch" if kind == "flop" else "algorithmic_bytes_per_launch"): work}

        per_kernel = {k: kernel_roofline(k) for k in model}
        dom = max(model, key=lambda k: stages[k])
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
        except Exception:
            pass
        roof = dict(per_kernel[dom])
        roof["traffic"] = traffic
        roof["peak_source"] = {
            "tensor": peak_src_t + ".  `achieved` counts the algorithmic conv flops (2*K1*F1*Chans*Samples per sample) "
                      "once; the kernel issues 3 tf32 MMAs per product (hi*hi + hi*lo + lo*hi keeps fp32 parity) and "
                      "its small-N tiles are bound by shared-memory operand reads, see DESIGN.md section 4.6",
            "fp32": "measured live: register-resident FFMA loop on all SMs (eav_measure_fp32_peak)",
            "hbm": "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if peaks else "fallback 7.7 TB/s nominal",
        }[roof["bound"]]
        roof["kernels"] = per_kernel
        roof["fp32_ceilings"] = {"ffma_pipe_peak": fp32_peak, "register_outer_product_ffma2": fp32_ffma2,
                                 "note": "CUDA-core ceilings on this device: immediate-operand FFMA loop, and an 8x8 "
                                         "register outer product with packed FFMA2 (the block-2 conv kernels' mix)"}
        roof["whole_step"] = {"achieved": FLOP_PER_SAMPLE * N / (ms * 1e-3) / 1e12, "unit": "TFLOP/s",
                              "note": "algorithmic conv/dense flops of fwd+bwd over the measured step time"}

    # ---------------------------------------------------------------- CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cms, threads = cpu_train_baseline(12, 2, bn_train=head_train)
        cpu = {"value": v, "unit": "samples/s", "cores": threads, "kind": "port",
               "sample": f"12 B=32 train steps ({args.bn_mode}-mode BN) of one subject model = 1/42 of a GPU step; "
                         "oracle/eegnet_oracle.py (torch-CPU restatement of CNN_torch/EEGNet_tor.py)",
               "ms_per_step": cms}
        if pre is not None:
            pv, pdt = cpu_preproc_baseline()
            pre["cpu_baseline"] = {"value": pv, "unit": "GB/s", "cores": 1, "kind": "port",
                                   "sample": f"1 of 42 subjects (264 MB algorithmic), oracle/preproc_oracle.c, {pdt:.2f} s"}

    if world > 1:
        dist.barrier(device_ids=[local])
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {"metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"{M} per-subject EEGNet_tor models per GPU trained in lock-step "
                                   "(BASELINE.json configs[2]): Chans=30 Samples=500 kernLength=300 F1=8 D=8 F2=64, "
                                   "5 classes, batch 32 per model, Adam lr=1e-5; synthetic 200x10000x30 EEG per subject "
                                   "preprocessed on the GPU, 280/120 split",
                       "subjects_per_gpu": M, "batch_per_model": B, "samples_per_step": M * B * world,
                       "bn_mode": args.bn_mode, "dropout": "on-device Philox" if head_train else "off (eval)",
                       "l2": "no explicit flush: a step streams ~3.5 GB of activations through a 126 MB L2, inputs "
                             "(80 MB batch gathered by index from a 0.7 GB resident set) exceed L2",
                       "parallelism": f"subject-sharded x{world}, no collective", "cuda_graph": True,
                       "arithmetic": "fp32 storage and accumulation everywhere; the temporal conv and the block-2 conv (fwd, "
                                     "dX, dW) run on tcgen05 as a 3-product tf32 split (hi*hi + hi*lo + lo*hi, ~2^-21 "
                                     "relative), all other kernels on the fp32 CUDA cores"
                                     if all(os.environ.get(k, "") not in ("ffma", "0")
                                            for k in ("EAV_TC", "EAV_TCONV", "EAV_SEPCONV")) else
                                     "tensor-core paths (partly) switched off by EAV_TC / EAV_TCONV / EAV_SEPCONV"},
            "gpu_launches": launches_per_step * K,
            "launches_per_step": launches_per_step,
            ("eval_bn_step" if head_train else "train_bn_step"): {"ms_per_step": other_ms,
                                                                   "value": world * M * B / (other_ms * 1e-3),
                                                                   "unit": "samples/s"},
            "final_mean_loss": loss_now,
            "e2e": {"value": e2e_value, "unit": "samples/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": M * 4,
                    "how": "pinned host batch -> cudaMemcpyAsync (double-buffered on a copy stream) -> "
                           "eav_eegnet_forward/loss/backward/adam through the C ABI -> loss D2H every step"},
            "clocks": clk}
    if roof is not None:
        line["roofline"] = roof
        line["stage_ms"] = stages
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if pre is not None:
        line["preprocess"] = pre
    _emit(line)


if __name__ == "__main__":
    main()
