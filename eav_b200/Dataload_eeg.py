"""Drop-in for the reference's Dataload_eeg.py `DataLoadEEG` (Dataload_eeg.py:35-160):
same constructor, attributes and method names; the filter / decimate / epoch arithmetic
runs on hand-written sm_100a kernels (libeav_b200.so: FIR decimation -> time-parallel SOS
scan -> epoch scatter).  `.mat` I/O and the two filter DESIGN calls (firwin taps as
scipy.signal.resample_poly builds them, butter(5, band, 'bandpass', 'sos')) stay on the host
exactly as in the reference.  No CPU fallback for the filtering itself.

prepare_data() is the fused fast path (raw -> epochs without round trips); the individual
stage methods are kept for scripts that call them one by one and leave the same attributes
behind (self.seg, self.seg_f, self.seg_f_div, self.label_div) as float32-valued arrays.
The README's legacy method names (data_mat / downsampling / bandpass / data_div /
data_prepare, used by EEGNet_tor.py:149) are provided as aliases.
"""
import os

import numpy as np
import scipy.io
import torch
from scipy.signal import butter, firwin

from .EAV_datasplit import EAVDataSplit  # noqa: F401  (re-exported like the reference's star import)
from .ops import PreprocEngine

SELECTED_CLASSES = [1, 3, 5, 7, 9]  # Classes corresponding to listening tasks (Dataload_eeg.py:33)


def decimation_taps(down):
    """The anti-aliasing FIR scipy.signal.resample_poly(x, 1, down) designs internally."""
    half_len = 10 * down
    return firwin(2 * half_len + 1, 1.0 / down, window=('kaiser', 5.0))


def epoch_slots(label, n_sub=4):
    """Integer plan of segment_and_select_classes (Dataload_eeg.py:139-152): per trial the output
    slot (-1 = dropped) and the labels of the kept epochs -- argmax over the 10 one-hot rows,
    so labels are in {1,3,5,7,9} exactly as the reference returns them."""
    cls = np.argmax(np.asarray(label), axis=0)
    keep = np.isin(cls, SELECTED_CLASSES)
    slot = np.full(cls.shape, -1, dtype=np.int32)
    slot[keep] = np.arange(int(keep.sum()), dtype=np.int32)
    return slot, np.repeat(cls[keep], n_sub).astype(np.int64)


class DataLoadEEG:
    """Loads and preprocesses the EEG of one subject (GPU-accelerated drop-in)."""

    def __init__(self, subject=1, band=[0.3, 50], fs_orig=500, fs_target=100,
                 parent_directory='./Datasets/EAV'):
        self.subject = subject
        self.band = band
        self.fs_orig = fs_orig
        self.fs_target = fs_target
        self.parent_directory = parent_directory

        self.seg = None        # (Channels, Time, Trials)
        self.label = None      # (10, Trials) one-hot
        self.seg_f = None      # filtered (Channels, Time, Trials)
        self.seg_f_div = None  # (Epochs, Channels, Samples)
        self.label_div = None  # (Epochs,)
        self.seg_f_div_device = None   # the same epochs, float32 on the GPU (N, Chans, Samples)
        self._raw_trial_major = None   # [trials][ch][time] view of the raw recording
        self._device = None

    # ------------------------------------------------------------------ I/O (host, as the reference)
    def load_mat_data(self):
        subject_str = f'subject{self.subject:02d}'
        eeg_folder = os.path.join(self.parent_directory, subject_str, 'EEG')
        base_name = subject_str.rstrip('__')
        eeg_file_path = os.path.join(eeg_folder, base_name + '_eeg.mat')
        label_file_path = os.path.join(eeg_folder, base_name + '_eeg_label.mat')
        if not os.path.exists(eeg_file_path):
            print(f'[Error] EEG data not found for {subject_str}')
            return
        # MAT v5 payload mapped in place when it is stored uncompressed (mat_ingest.read_mat_array); scipy.io.loadmat
        # is the fallback for anything else.  `raw` is [trial][ch][time] = the file's own memory order.
        from .mat_ingest import load_subject_mat
        raw, label, _ = load_subject_mat(self.parent_directory, self.subject)
        self.label = label
        self.seg = np.transpose(raw, (1, 2, 0))                      # (Channels, Time, Trials) view, no copy
        self._raw_trial_major = raw
        print(f'[Info] Loaded EEG data for {subject_str}')

    def set_raw(self, cnt, label):
        """cnt: (Time, Channels, Trials) as stored in the .mat (Dataload_eeg.py:81-82)."""
        self.label = np.asarray(label)
        self.seg = np.transpose(cnt, [1, 0, 2])                     # (Channels, Time, Trials)
        self._raw_trial_major = None

    # ------------------------------------------------------------------ helpers
    def _dev(self):
        if self._device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("eav_b200.DataLoadEEG needs a CUDA (B200) device; there is no CPU fallback")
            self._device = torch.device("cuda", torch.cuda.current_device())
        return self._device

    def _raw_device(self, seg):
        """(Channels, Time, Trials) host array -> device tensor [1][trials][ch][time] (the .mat memory order)."""
        # the cached trial-major view is only trusted while self.seg still IS a view of it (a caller may have
        # re-assigned loader.seg after load_mat_data())
        if (seg is self.seg and self._raw_trial_major is not None and isinstance(seg, np.ndarray)
                and seg.shape == tuple(np.transpose(self._raw_trial_major, (1, 2, 0)).shape)
                and np.shares_memory(seg, self._raw_trial_major)):
            arr = np.ascontiguousarray(self._raw_trial_major)          # already in the kernels' layout
        else:
            arr = np.ascontiguousarray(np.transpose(np.asarray(seg), (2, 0, 1)))
        if arr.dtype not in (np.float32, np.float64):
            arr = arr.astype(np.float64)
        return torch.from_numpy(arr).unsqueeze(0).to(self._dev())

    def _engine(self, ch, t, tri, dtype, n_sub):
        down = int(self.fs_orig / self.fs_target)
        return PreprocEngine(1, n_trials=tri, n_chans=ch, trial_len=t, down=down, n_taps=2 * 10 * down + 1,
                             n_sections=5, n_sub=n_sub, raw_dtype=dtype, device=self._dev()), down

    def _sos(self):
        return butter(5, self.band, btype='bandpass', fs=self.fs_target, output='sos')   # Dataload_eeg.py:113

    # ------------------------------------------------------------------ staged API (reference method names)
    def downsampling(self):
        if self.seg is None:
            return
        ch, t, tri = self.seg.shape
        raw = self._raw_device(self.seg)
        eng, down = self._engine(ch, t, tri, raw.dtype, 1)
        slot = torch.full((1, tri), -1, dtype=torch.int32, device=self._dev())
        # identity SOS: only the decimated sequence is wanted from this call
        ident = np.tile(np.array([1.0, 0, 0, 1, 0, 0]), (5, 1))
        _, dec = eng.run(raw, decimation_taps(down), ident, slot, 0, want_dec=True)
        new_time = int(t * (self.fs_target / self.fs_orig))
        self._raw_trial_major = None                                        # self.seg stops being the raw recording
        self._dec_device = dec                                              # [1][ch][tri*new_time]
        self.seg = dec[0].reshape(ch, tri, new_time).permute(0, 2, 1).cpu().numpy()   # (ch, new_time, tri)

    def bandpass_filter(self):
        if self.seg is None:
            return
        ch, t, tri = self.seg.shape
        # the decimated sequence re-enters the same engine with a unit "decimation" (1 tap, down=1)
        raw = self._raw_device(self.seg)
        eng = PreprocEngine(1, n_trials=tri, n_chans=ch, trial_len=t, down=1, n_taps=1, n_sections=5, n_sub=1,
                            raw_dtype=raw.dtype, device=self._dev())
        slot = torch.arange(tri, dtype=torch.int32, device=self._dev()).unsqueeze(0)
        ep = eng.run(raw, np.ones(1), self._sos(), slot, tri)              # [1][tri][ch][t]
        self.seg_f = ep[0].permute(1, 2, 0).cpu().numpy()                   # (ch, t, tri)

    def segment_and_select_classes(self):
        if self.seg_f is None:
            return
        # index logic on the host, the reference's own shapes (30/500/4/200 hard-coded, Dataload_eeg.py:133-136)
        tm1 = self.seg_f.reshape((30, 500, 4, 200), order='F')
        seg = tm1.reshape((30, 500, 4 * 200), order='F')
        label_div = np.repeat(self.label, repeats=4, axis=1)
        selected_mask = np.isin(np.argmax(label_div, axis=0), SELECTED_CLASSES)
        self.seg_f_div = np.transpose(seg[:, :, selected_mask], (2, 0, 1))
        self.label_div = np.argmax(label_div[:, selected_mask], axis=0)
        self.seg_f_div_device = None

    # ------------------------------------------------------------------ fused fast path
    def prepare_data(self):
        """load -> decimate -> band-pass -> epoch/select in one GPU pipeline.  Returns
        (x float32 (N, Chans, Samples) numpy, y int64 (N,) numpy) like the reference
        (Dataload_eeg.py:154-160); the device copy stays in self.seg_f_div_device."""
        self.load_mat_data()
        if self.seg is None:
            return self.seg_f_div, self.label_div
        x_dev, y = self.prepare_data_device()
        self.seg_f_div = x_dev.cpu().numpy()
        self.label_div = y
        return self.seg_f_div, self.label_div

    def prepare_data_device(self):
        """Same as prepare_data() without the .mat load and without leaving the GPU."""
        ch, t, tri = self.seg.shape
        raw = self._raw_device(self.seg)
        n_sub = 4
        eng, down = self._engine(ch, t, tri, raw.dtype, n_sub)
        slot, y = epoch_slots(self.label, n_sub)
        n_ep = int((slot >= 0).sum()) * n_sub
        ep = eng.run(raw, decimation_taps(down), self._sos(),
                     torch.from_numpy(slot).unsqueeze(0).to(self._dev()), n_ep)
        self.seg_f_div_device = ep[0]
        self.label_div = y
        return self.seg_f_div_device, y

    # ------------------------------------------------------------------ legacy order (the paper's 5-class setting)
    def prepare_data_legacy_device(self, band=(3, 50)):
        """CNN_tensorflow/CNN_EEG_tf.py:64-75,180-206 on the GPU: band-pass the raw recording at fs_orig
        (butter(5, band, 'band', fs=fs_orig), continuous over the trials), THEN resample_poly to fs_target, split
        every trial into 4 epochs, keep classes {1,3,5,7,9} and remap them to 0..4 (rows [1,3,5,7,9] of the one-hot
        label matrix, :206).  Returns (x [N][Chans][500] float32 on the device, y int64 (N,) in 0..4)."""
        ch, t, tri = self.seg.shape
        seg = self.seg if np.asarray(self.seg).dtype == np.float32 else np.asarray(self.seg, dtype=np.float32)
        raw = self._raw_device(seg)
        n_sub = 4
        down = int(self.fs_orig / self.fs_target)
        eng = PreprocEngine(1, n_trials=tri, n_chans=ch, trial_len=t, down=down, n_taps=2 * 10 * down + 1,
                            n_sections=5, n_sub=n_sub, raw_dtype=torch.float32, device=self._dev(), order=1)
        slot, y = epoch_slots(self.label, n_sub)
        n_ep = int((slot >= 0).sum()) * n_sub
        sos = butter(5, list(band), btype='band', fs=self.fs_orig, output='sos')       # CNN_EEG_tf.py:69
        ep = eng.run(raw, decimation_taps(down), sos, torch.from_numpy(slot).unsqueeze(0).to(self._dev()), n_ep)
        self.seg_f_div_device = ep[0]
        self.label_div = (np.asarray(y) - 1) // 2
        return self.seg_f_div_device, self.label_div

    def prepare_data_legacy(self, band=(3, 50)):
        """load_mat_data() + prepare_data_legacy_device(); returns numpy (x float32 (N, Chans, 500), y in 0..4),
        which EAVDataSplit(x, y).get_split(h_idx=56) splits 280/120 as the paper does."""
        self.load_mat_data()
        if self.seg is None:
            return self.seg_f_div, self.label_div
        x_dev, y = self.prepare_data_legacy_device(band)
        self.seg_f_div = x_dev.cpu().numpy()
        return self.seg_f_div, self.label_div

    # legacy names documented in the reference README (README.md:186-199) and used at EEGNet_tor.py:149
    data_mat = load_mat_data
    bandpass = bandpass_filter
    data_div = segment_and_select_classes
    data_prepare = prepare_data
