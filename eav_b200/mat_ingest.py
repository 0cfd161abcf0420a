"""`.mat` ingest -> device (replaces the I/O half of DataLoadEEG.load_mat_data, Dataload_eeg.py:54-83).

The reference calls scipy.io.loadmat per subject (1.4 s for the 480 MB float64 `seg`), which parses, copies and
transposes on one core.  Once filtering and training take milliseconds that read is the wall-clock bound, so:

* `read_mat_array` parses the MAT-file v5 container itself and returns the array WITHOUT a copy when the element is
  stored uncompressed (np.memmap straight onto the payload), or inflates a compressed element in one pass.  The
  column-major (Time, Channels, Trials) payload is byte-for-byte the [trial][ch][time] layout the kernels take.
* `SubjectPrefetcher` reads subject k+1 on a background thread, converts to float32 into a pinned staging buffer
  and issues the H2D copy on its own CUDA stream while the GPU still works on subject k.

Anything the parser does not understand (v7.3/HDF5, sparse, complex, cell arrays) falls back to scipy.io.loadmat.
"""
import os
import queue
import struct
import threading
import zlib

import numpy as np

# MAT-file v5 data types (MAT-File Format, table 1-1) -> numpy
_MI = {1: "i1", 2: "u1", 3: "i2", 4: "u2", 5: "i4", 6: "u4", 7: "f4", 9: "f8", 12: "i8", 13: "u8"}
_MI_MATRIX, _MI_COMPRESSED = 14, 15
_NUMERIC_CLASSES = {6, 7, 8, 9, 10, 11, 12, 13, 14, 15}   # mxDOUBLE .. mxUINT64


class MatFormatError(ValueError):
    pass


def _tag(buf, off, end):
    """(type, nbytes, data offset, offset of the next element) of the data element at `off`."""
    if off + 8 > len(buf):
        raise MatFormatError("truncated element tag")
    w0, w1 = struct.unpack_from(end + "II", buf, off)
    if w0 >> 16:                                   # small data element: 2-byte size, 2-byte type, 4 data bytes
        return w0 & 0xFFFF, w0 >> 16, off + 4, off + 8
    nxt = off + 8 + w1
    if w0 != _MI_COMPRESSED:
        nxt = (nxt + 7) & ~7
    return w0, w1, off + 8, nxt


def _parse_matrix(buf, off, nbytes, end):
    """Header of one miMATRIX element: (name, dims, numpy dtype, payload offset, payload bytes) or None."""
    stop = off + nbytes
    t, n, d, off = _tag(buf, off, end)             # array flags
    if n < 8:
        return None
    flags = struct.unpack_from(end + "I", buf, d)[0]
    klass, is_complex = flags & 0xFF, bool(flags & 0x0800)
    t, n, d, off = _tag(buf, off, end)             # dimensions
    dims = struct.unpack_from(end + "%di" % (n // 4), buf, d)
    t, n, d, off = _tag(buf, off, end)             # name
    name = bytes(buf[d:d + n]).decode("latin1")
    if klass not in _NUMERIC_CLASSES or is_complex or off >= stop:
        return name, dims, None, 0, 0
    t, n, d, off = _tag(buf, off, end)             # real part
    if t not in _MI:
        return name, dims, None, 0, 0
    # MATLAB may store a class-double array in a narrower element type ("numeric data compression").  The reference
    # calls scipy.io.loadmat with its defaults (Dataload_eeg.py:70,77; mat_dtype=False), which returns the STORAGE
    # dtype in that case -- so does this reader (pinned by tests/test_mat_ingest_cpu.py against loadmat itself);
    # DataLoadEEG widens non-float recordings to float64 before they reach the kernels, as scipy's filters would.
    return name, dims, np.dtype(end + _MI[t]), d, n


def read_mat_array(path, names):
    """First array among `names` found in the MAT-file v5 at `path`, as a C-ordered view of the REVERSED MATLAB
    dims: a (10000, 30, 200) `seg` comes back as [200][30][10000] (trial, channel, time), i.e. arr[k, c, t] ==
    loadmat(path)[name][t, c, k].  Returns (array, name, zero_copy: bool)."""
    with open(path, "rb") as f:
        head = f.read(128)
    if len(head) < 128 or head[:4] == b"\x89HDF" or head[:10] != b"MATLAB 5.0":
        raise MatFormatError("not a MAT-file v5 (v7.3 files are HDF5)")
    end = "<" if head[126:128] == b"IM" else ">"
    mm = np.memmap(path, dtype=np.uint8, mode="r")
    off = 128
    while off + 8 <= mm.shape[0]:
        t, n, d, nxt = _tag(mm, off, end)
        if t == _MI_MATRIX:
            info = _parse_matrix(mm, d, n, end)
            if info and info[0] in names and info[2] is not None:
                name, dims, dt, po, pn = info
                count = int(np.prod(dims))
                if pn < count * dt.itemsize:
                    raise MatFormatError("truncated payload")
                arr = np.memmap(path, dtype=dt, mode="r", offset=po, shape=tuple(reversed(dims)))
                return arr, name, True
        elif t == _MI_COMPRESSED:
            # peek at the header of the inflated element before paying for the whole stream
            dec = zlib.decompressobj()
            first = dec.decompress(bytes(mm[d:d + min(n, 4096)]), 512)
            it, inb, idat, _ = _tag(first, 0, end)
            if it == _MI_MATRIX:
                try:
                    info = _parse_matrix(first, idat, min(inb, len(first) - idat), end)
                except (MatFormatError, struct.error):
                    info = None
                if info and info[0] in names and info[2] is not None:
                    name, dims, dt, po, pn = info
                    raw = zlib.decompress(bytes(mm[d:d + n]))
                    count = int(np.prod(dims))
                    if pn < count * dt.itemsize:
                        raise MatFormatError("truncated payload")
                    arr = np.frombuffer(raw, dtype=dt, count=count, offset=po).reshape(tuple(reversed(dims)))
                    return arr, name, False
        off = nxt
    raise KeyError(f"none of {names} in {path}")


def load_subject_mat(parent_directory, subject):
    """(raw [trial][ch][time] in the stored dtype, label (10, trials) float64, zero_copy) for one subject; the
    file layout and the `seg1`-before-`seg` preference are the reference's (Dataload_eeg.py:56-79)."""
    subject_str = f"subject{subject:02d}"
    folder = os.path.join(parent_directory, subject_str, "EEG")
    eeg, lab = os.path.join(folder, subject_str + "_eeg.mat"), os.path.join(folder, subject_str + "_eeg_label.mat")
    if not os.path.exists(eeg):
        raise FileNotFoundError(eeg)
    import scipy.io
    try:
        try:
            raw, _, zero_copy = read_mat_array(eeg, ("seg1",))
        except KeyError:
            raw, _, zero_copy = read_mat_array(eeg, ("seg",))
    except (MatFormatError, KeyError, struct.error, zlib.error):
        mat = scipy.io.loadmat(eeg)
        cnt = np.array(mat.get("seg1")) if "seg1" in mat else np.array(mat.get("seg"))
        raw, zero_copy = np.ascontiguousarray(np.transpose(cnt, (2, 1, 0))), False
    label = np.array(scipy.io.loadmat(lab).get("label"))          # 16 KB: not worth a custom path
    return raw, label, zero_copy


class SubjectPrefetcher:
    """Iterates (subject, raw_device [trial][ch][time] float32, label) with the next subject's file read, float32
    conversion and host->device copy overlapped with the caller's GPU work on the current one.

    depth staging slots (pinned host buffer + device buffer each) rotate; a slot is reused only after the
    consumer asked for the next item, so keep at most `depth - 1` yielded tensors alive.  device=None keeps
    everything on the host (CPU tests)."""

    def __init__(self, parent_directory, subjects, device="cuda", depth=2, chunk_trials=8):
        self.parent, self.subjects, self.depth, self.chunk = parent_directory, list(subjects), max(2, depth), chunk_trials
        self.device = device
        self._q = queue.Queue(maxsize=self.depth - 1)
        self._slots = []
        self._free = queue.Queue()
        self._err = None
        self._thread = None
        self.read_seconds = 0.0

    def _slot(self, shape):
        import torch
        if self.device is None:
            return {"host": torch.empty(shape, dtype=torch.float32), "dev": None, "event": None, "done": None}
        host = torch.empty(shape, dtype=torch.float32, pin_memory=True)
        return {"host": host, "dev": torch.empty(shape, dtype=torch.float32, device=self.device),
                "event": torch.cuda.Event(), "done": None}

    def _worker(self):
        import time
        import torch
        try:
            stream = torch.cuda.Stream(device=self.device) if self.device is not None else None
            for s in self.subjects:
                t0 = time.perf_counter()
                raw, label, _ = load_subject_mat(self.parent, s)
                if len(self._slots) < self.depth:
                    slot = self._slot(tuple(raw.shape))
                    self._slots.append(slot)
                else:
                    slot = self._free.get()
                    if slot is None:
                        return
                if slot["event"] is not None and slot.get("done") is not None:
                    slot["event"].synchronize()        # the previous H2D out of this pinned buffer has finished
                if tuple(slot["host"].shape) != tuple(raw.shape):
                    raise ValueError(f"subject {s}: recording shape {raw.shape} differs from the first subject's")
                hnp = slot["host"].numpy()
                for k in range(0, raw.shape[0], self.chunk):          # page-in + narrow to float32, chunk by chunk
                    np.copyto(hnp[k:k + self.chunk], raw[k:k + self.chunk], casting="same_kind")
                self.read_seconds += time.perf_counter() - t0
                if stream is not None:
                    with torch.cuda.stream(stream):
                        if slot["done"] is not None:   # the consumer's kernels on the old contents have finished
                            stream.wait_event(slot["done"])
                        slot["dev"].copy_(slot["host"], non_blocking=True)
                        slot["event"].record(stream)
                self._q.put((s, slot, label))
            self._q.put(None)
        except BaseException as e:  # noqa: BLE001
            self._err = e
            self._q.put(None)

    def __iter__(self):
        import torch
        self._thread = threading.Thread(target=self._worker, daemon=True)
        self._thread.start()
        prev = None
        while True:
            item = self._q.get()
            if prev is not None:               # the consumer moved on: its queued GPU work on the slot gates the reuse
                if prev["event"] is not None:
                    prev["done"] = torch.cuda.Event()
                    prev["done"].record(torch.cuda.current_stream(self.device))
                self._free.put(prev)
                prev = None
            if item is None:
                break
            s, slot, label = item
            if slot["event"] is not None:
                torch.cuda.current_stream(self.device).wait_event(slot["event"])
                out = slot["dev"]
            else:
                out = slot["host"]
            prev = slot
            yield s, out, label
        self._free.put(None)
        if self._err is not None:
            raise self._err


def prepare_subjects(parent_directory, subjects, band=(0.3, 50), fs_orig=500, fs_target=100, device="cuda",
                     legacy_order=False):
    """DataLoadEEG(...).prepare_data() for a list of subjects with the file read / H2D of the next subject
    overlapped with the filtering of the current one.  Yields (subject, epochs [N][Chans][500] float32 on the
    device, y int64 numpy) -- y in {1,3,5,7,9} as the reference returns it, or 0..4 with
    legacy_order=True (band-pass at fs_orig first, CNN_EEG_tf.py:180-206)."""
    import torch
    from scipy.signal import butter
    from .Dataload_eeg import decimation_taps, epoch_slots
    from .ops import PreprocEngine
    down = int(fs_orig / fs_target)
    eng = None
    for s, raw, label in SubjectPrefetcher(parent_directory, subjects, device=device):
        tri, ch, t = raw.shape
        if eng is None:
            eng = PreprocEngine(1, n_trials=tri, n_chans=ch, trial_len=t, down=down, n_taps=2 * 10 * down + 1,
                                n_sections=5, n_sub=4, raw_dtype=torch.float32, device=device,
                                order=1 if legacy_order else 0)
        slot, y = epoch_slots(label, 4)
        n_ep = int((slot >= 0).sum()) * 4
        sos = butter(5, list(band), btype="bandpass", fs=fs_orig if legacy_order else fs_target, output="sos")
        ep = eng.run(raw.unsqueeze(0), decimation_taps(down), sos, torch.from_numpy(slot).unsqueeze(0).to(device), n_ep)
        yield s, ep[0], ((np.asarray(y) - 1) // 2 if legacy_order else np.asarray(y))
