"""CPU oracle for the EAV EEG hot path  (TEST INFRASTRUCTURE -- never shipped).

A restatement, in numpy / torch-CPU / plain C (preproc_oracle.c), of what the
reference (nubcico/EAV) computes on the path
    Dataload_eeg.py filter/decimate/epoch -> EAV_datasplit.py split ->
    CNN_torch/EEGNet_tor.py (and CNN_torch/CNN_EEG.py) forward/backward/Adam.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this package, and only as the checker or the
timed CPU baseline.  The product (eav_b200/) never imports it and has no CPU
fallback.

Parity status: PINNED.  The reference ships no tests or golden vectors (SURVEY.md
section 4), so the pins are outputs of the unmodified reference executed in the
build container through oracle/ref_shim.py; oracle/gen_golden.py writes them to
tests/golden/*.npz and tests/test_oracle_golden.py checks every function here
against them (plus the known-answer constants of SURVEY.md section 8c).

Third-party arithmetic on the path (absent from /root/reference, unpinned in its
requirements.txt; versions are the image's): scipy 1.18.1 `signal.resample_poly`
(firwin + upfirdn), `signal.butter(output='sos')`, `signal.sosfilt`; numpy 2.3.5
indexing; torch 2.11.0 ATen CPU kernels.  Their published algorithms are restated
here; the filter DESIGN (`butter`) is delegated to scipy on both sides exactly as
the reference calls it (Dataload_eeg.py:113) and pinned by a known-answer test.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_LIB = None

SELECTED_CLASSES = (1, 3, 5, 7, 9)  # Dataload_eeg.py:33


# ----------------------------------------------------------------------------
# C restatement loader
# ----------------------------------------------------------------------------
def build_c(force: bool = False) -> str:
    """gcc-compile preproc_oracle.c into oracle/_build/liboracle_preproc.so."""
    os.makedirs(_BUILD, exist_ok=True)
    src = os.path.join(_HERE, "preproc_oracle.c")
    out = os.path.join(_BUILD, "liboracle_preproc.so")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", out, src])
    return out


def _lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build_c())
        i64, dp, fp, ci = ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int
        L.oracle_fir_decimate.argtypes = [dp, dp, ci, ci, dp, i64]
        L.oracle_sosfilt.argtypes = [dp, ci, dp, i64]
        L.oracle_epoch_gather.argtypes = [dp, ci, i64, i64, ci, dp, dp]
        L.oracle_epoch_gather.restype = i64
        _LIB = L
    return _LIB


def _ptr(a: np.ndarray):
    return ctypes.c_void_p(a.ctypes.data)


# ----------------------------------------------------------------------------
# P2  decimating FIR  (Dataload_eeg.py:85-102 -> scipy.signal.resample_poly)
# ----------------------------------------------------------------------------
def kaiser_window(n: int, beta: float) -> np.ndarray:
    """scipy.signal.windows.kaiser(n, beta, sym=True): I0(beta*sqrt(1-(2k/(n-1)-1)^2))/I0(beta)."""
    k = np.arange(n, dtype=np.float64)
    alpha = (n - 1) / 2.0
    return np.i0(beta * np.sqrt(np.maximum(0.0, 1.0 - ((k - alpha) / alpha) ** 2))) / np.i0(beta)


def decimation_taps(down: int) -> np.ndarray:
    """The low-pass resample_poly(up=1, down) designs: firwin(2*10*down+1, 1/down,
    window=('kaiser', 5.0)) scaled by up=1.  firwin's published algorithm for a
    single-band low-pass: h[n] = fc*sinc(fc*(n-alpha)) * w[n], then divided by its
    DC gain so that sum(h) == 1."""
    half_len = 10 * down
    numtaps = 2 * half_len + 1
    fc = 1.0 / down
    m = np.arange(numtaps, dtype=np.float64) - half_len
    h = fc * np.sinc(fc * m) * kaiser_window(numtaps, 5.0)
    return h / h.sum()


def fir_decimate(raw: np.ndarray, taps: np.ndarray, down: int) -> np.ndarray:
    """raw: [trials][ch][trial_len] (the .mat memory order), f32 or f64.
    Returns the decimated CONTINUOUS sequences [ch][trials*trial_len/down] f64:
    out[c, j] = sum_d h[H+d] * x_c[down*j - d], zero outside the record."""
    raw = np.asarray(raw)
    n_tr, n_ch, tl = raw.shape
    n = n_tr * tl
    n_out = -(-n // down)
    taps = np.ascontiguousarray(taps, dtype=np.float64)
    H = (taps.size - 1) // 2
    out = np.empty((n_ch, n_out), dtype=np.float64)
    seq = np.zeros(n + 2 * H + down, dtype=np.float64)
    L = _lib()
    for c in range(n_ch):
        # order='F' reshape of (ch, t, trials) at Dataload_eeg.py:94 == trials concatenated in time
        seq[H:H + n] = raw[:, c, :].reshape(-1)
        L.oracle_fir_decimate(_ptr(seq), _ptr(taps), H, down, _ptr(out[c]), n_out)
    return out


def fir_decimate_numpy(x: np.ndarray, taps: np.ndarray, down: int) -> np.ndarray:
    """Same closed form, pure numpy, for one continuous sequence x[n] (small cases)."""
    x = np.asarray(x, dtype=np.float64)
    H = (taps.size - 1) // 2
    n_out = -(-x.size // down)
    xp = np.concatenate([np.zeros(H), x, np.zeros(H + down)])
    out = np.zeros(n_out)
    j = np.arange(n_out)
    for d in range(-H, H + 1):
        out += taps[H + d] * xp[H + down * j - d]
    return out


# ----------------------------------------------------------------------------
# P3  band-pass  (Dataload_eeg.py:104-121 -> butter(..., 'sos') + sosfilt)
# ----------------------------------------------------------------------------
def butter_sos(band, fs: float, order: int = 5) -> np.ndarray:
    """Filter DESIGN is delegated to scipy exactly as the reference calls it
    (Dataload_eeg.py:113); pinned by the known-answer SOS of SURVEY.md 8c."""
    from scipy.signal import butter
    return np.ascontiguousarray(butter(order, band, btype="bandpass", fs=fs, output="sos"), dtype=np.float64)


def sosfilt(sos: np.ndarray, seqs: np.ndarray) -> np.ndarray:
    """Cascaded DF2T biquads, zero initial state, per row of seqs [ch][n] (f64)."""
    sos = np.ascontiguousarray(sos, dtype=np.float64)
    out = np.array(seqs, dtype=np.float64, order="C", copy=True)
    L = _lib()
    for c in range(out.shape[0]):
        L.oracle_sosfilt(_ptr(sos), sos.shape[0], _ptr(out[c]), out.shape[1])
    return out


def sosfilt_python(sos: np.ndarray, x: np.ndarray) -> np.ndarray:
    """Pure-Python DF2T loop (small cases only) -- same recurrence as the C one."""
    y = np.array(x, dtype=np.float64)
    zi = np.zeros((sos.shape[0], 2))
    for i in range(y.size):
        xc = y[i]
        for s in range(sos.shape[0]):
            b0, b1, b2, _, a1, a2 = sos[s]
            xn = b0 * xc + zi[s, 0]
            zi[s, 0] = b1 * xc - a1 * xn + zi[s, 1]
            zi[s, 1] = b2 * xc - a2 * xn
            xc = xn
        y[i] = xc
    return y


# ----------------------------------------------------------------------------
# P4  epoching + class selection  (Dataload_eeg.py:123-152)
# ----------------------------------------------------------------------------
def trial_classes(label: np.ndarray) -> np.ndarray:
    """argmax over the 10 one-hot rows (label is (10, n_trials))."""
    return np.argmax(label, axis=0)


def epoch_plan(label: np.ndarray, n_sub: int = 4):
    """Integer plan of segment_and_select_classes: which trials are kept, the label
    of every kept epoch (values in {1,3,5,7,9}: SURVEY F7), and for every kept epoch
    its (trial, sub-epoch) source.  Epoch e = n_sub*k + q <- trial k, samples
    [q*ep_len, (q+1)*ep_len)."""
    cls = trial_classes(label)
    keep = np.isin(cls, SELECTED_CLASSES)
    kept_trials = np.nonzero(keep)[0]
    src_trial = np.repeat(kept_trials, n_sub)
    src_sub = np.tile(np.arange(n_sub), kept_trials.size)
    y = np.repeat(cls[kept_trials], n_sub).astype(np.int64)
    return keep, y, src_trial, src_sub


def epoch_gather(seqs: np.ndarray, label: np.ndarray, trial_len: int, n_sub: int = 4):
    """seqs [ch][n_trials*trial_len] -> (x [n_epochs][ch][trial_len/n_sub] f64, y int64)."""
    seqs = np.ascontiguousarray(seqs, dtype=np.float64)
    keep, y, _, _ = epoch_plan(label, n_sub)
    n_ch = seqs.shape[0]
    n_tr = seqs.shape[1] // trial_len
    out = np.empty((int(keep.sum()) * n_sub, n_ch, trial_len // n_sub), dtype=np.float64)
    keep8 = np.ascontiguousarray(keep.astype(np.uint8))
    n = _lib().oracle_epoch_gather(_ptr(seqs), n_ch, n_tr, trial_len, n_sub, _ptr(keep8), _ptr(out))
    assert n == out.shape[0]
    return out, y


def prepare_data(raw: np.ndarray, label: np.ndarray, band, fs_orig=500, fs_target=100):
    """Dataload_eeg.py:154-160 minus the .mat I/O.  raw: [trials][ch][time]."""
    down = int(fs_orig / fs_target)
    dec = fir_decimate(raw, decimation_taps(down), down)
    filt = sosfilt(butter_sos(band, fs_target), dec)
    trial_len = int(raw.shape[2] * (fs_target / fs_orig))
    return epoch_gather(filt, label, trial_len, 4)


def prepare_data_legacy(raw: np.ndarray, label: np.ndarray, band=(3, 50), fs_orig=500, fs_target=100):
    """The legacy pipeline of CNN_tensorflow/CNN_EEG_tf.py (the paper's 5-class 280/120 setting), minus .mat I/O:
      :64-75   Bandpass: butter(5, band, 'band', fs=fs_orig) + sosfilt on the continuous 500 Hz sequence of each
               channel (trials concatenated in time), i.e. the band-pass runs BEFORE the decimation;
      :182-189 resample_poly(up=1, down=fs_orig/fs_target) of that sequence;
      :84-101  mysplit: trial k -> epochs 4k..4k+3 (500 samples each), labels repeated 4x;
      :191-206 keep the trials of classes {1,3,5,7,9}; label_5class = rows [1,3,5,7,9] of the one-hot matrix,
               i.e. class index (c - 1) / 2 in 0..4.
    raw: [trials][ch][time].  Returns (x [n_epochs][ch][500] f64, y int64 in 0..4, onehot (5, n_epochs))."""
    down = int(fs_orig / fs_target)
    raw = np.asarray(raw)
    n_tr, n_ch, tl = raw.shape
    seqs = np.ascontiguousarray(np.transpose(raw, (1, 0, 2)).reshape(n_ch, n_tr * tl), dtype=np.float64)
    filt = sosfilt(butter_sos(band, fs_orig), seqs)                       # [ch][trials*tl] at fs_orig
    dec = fir_decimate(filt.reshape(n_ch, n_tr, tl).transpose(1, 0, 2), decimation_taps(down), down)
    x, y = epoch_gather(dec, label, tl // down, 4)
    y5 = (y - 1) // 2
    onehot = (np.arange(5)[:, None] == y5[None, :]).astype(np.float64)
    return x, y5, onehot


# ----------------------------------------------------------------------------
# S1-S3  split  (EAV_datasplit.py:12-40)
# ----------------------------------------------------------------------------
def split_indices(y: np.ndarray, h_idx: int = 40, n_classes: int = 5):
    """Indices into (x, y) the reference's get_split selects: class-major,
    original order within a class, first h_idx -> train, rest -> test; labels
    outside range(5) are silently dropped."""
    y = np.asarray(y)
    tr, te = [], []
    for cls in range(n_classes):
        idx = np.nonzero(y == cls)[0]
        tr.append(idx[:h_idx])
        te.append(idx[h_idx:])
    return np.concatenate(tr).astype(np.int64), np.concatenate(te).astype(np.int64)


def get_split(x: np.ndarray, y: np.ndarray, h_idx: int = 40):
    x = np.asarray(x)
    y = np.asarray(y)
    tr, te = split_indices(y, h_idx)
    return np.squeeze(x[tr]), y[tr], np.squeeze(x[te]), y[te]


# ----------------------------------------------------------------------------
# Synthetic dataset-shaped inputs  (SURVEY.md 8d / BASELINE.md section 3)
# ----------------------------------------------------------------------------
def synth_subject(subject: int, n_trials: int = 200, n_ch: int = 30, trial_len: int = 10000,
                  fs: float = 500.0):
    """raw EEG float32 [trials][ch][time] = N(0,1) + 0.5*sin(2*pi*50Hz*t) + 5*sin(2*pi*0.1Hz*t)
    (t continuous across trials), labels one-hot (10, n_trials) with n_trials/10 per
    class in rng.permutation order.  rng = default_rng(1000 + subject)."""
    rng = np.random.default_rng(1000 + subject)
    raw = rng.standard_normal((n_trials, n_ch, trial_len), dtype=np.float32)
    t = (np.arange(n_trials * trial_len, dtype=np.float64) / fs).reshape(n_trials, 1, trial_len)
    raw += (0.5 * np.sin(2 * np.pi * 50.0 * t) + 5.0 * np.sin(2 * np.pi * 0.1 * t)).astype(np.float32)
    cls = rng.permutation(np.repeat(np.arange(10), n_trials // 10))
    label = np.zeros((10, n_trials), dtype=np.float64)
    label[cls, np.arange(n_trials)] = 1.0
    return raw, label
