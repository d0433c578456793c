"""ctypes binding of include/sfgpu.h (libsfgpu.so).  Used by the tests and bench.py; the C host
(sigfish_b200/host) links the library directly.  There is no fallback: a missing library or a
missing GPU is an error."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
# SFGPU_LIB: another build of the same ABI (A/B experiments); the default is the in-tree library
LIB_PATH = os.environ.get("SFGPU_LIB") or os.path.join(PKG, "libsfgpu.so")

SFGPU_RNA, SFGPU_DTW, SFGPU_INV, SFGPU_REF, SFGPU_END, SFGPU_SAM = 0x001, 0x002, 0x004, 0x010, 0x020, 0x100

# every symbol include/sfgpu.h declares
SYMBOLS = ["sfgpu_device_count", "sfgpu_create", "sfgpu_set_ref", "sfgpu_submit", "sfgpu_resubmit",
           "sfgpu_collect", "sfgpu_timing", "sfgpu_destroy", "sfgpu_strerror", "sfgpu_ref_events",
           "sfgpu_event_table", "sfgpu_query", "sfgpu_ref_columns", "sfgpu_set_ref_events",
           "sfgpu_submit_queries", "sfgpu_collect_paths", "sfgpu_wave_reads",
           "sfgpu_submit_reads", "sfgpu_submit_records", "sfgpu_slot_signal"]


class Opt(C.Structure):
    _fields_ = [("device", C.c_int32), ("flags", C.c_uint32), ("query_size", C.c_int32),
                ("prefix_size", C.c_int32), ("kmer_size", C.c_int32), ("n_slots", C.c_int32),
                ("pore", C.c_int32), ("reserved", C.c_int32 * 5)]


class Result(C.Structure):
    _fields_ = [("n_events", C.c_int64), ("qstart", C.c_int32), ("qend", C.c_int32), ("qlen", C.c_int32),
                ("status", C.c_int32), ("start_raw", C.c_uint64), ("end_raw", C.c_uint64),
                ("score", C.c_float), ("score2", C.c_float), ("rid", C.c_int32), ("strand", C.c_int32),
                ("pos_st", C.c_int32), ("pos_end", C.c_int32)]


RESULT_DTYPE = np.dtype([("n_events", "<i8"), ("qstart", "<i4"), ("qend", "<i4"), ("qlen", "<i4"),
                         ("status", "<i4"), ("start_raw", "<u8"), ("end_raw", "<u8"), ("score", "<f4"),
                         ("score2", "<f4"), ("rid", "<i4"), ("strand", "<i4"), ("pos_st", "<i4"),
                         ("pos_end", "<i4")], align=True)
assert RESULT_DTYPE.itemsize == C.sizeof(Result)


class Timing(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("events_ms", C.c_float), ("dtw_ms", C.c_float),
                ("trace_ms", C.c_float), ("d2h_ms", C.c_float), ("total_ms", C.c_float),
                ("cells", C.c_double), ("samples", C.c_int64), ("dtw_launches", C.c_int32),
                ("other_launches", C.c_int32), ("tasks_per_read", C.c_int32), ("piece_blocks", C.c_int32),
                ("redone_pieces", C.c_int32), ("pad", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -m sigfish_b200.build` (no CPU fallback exists)")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.sfgpu_device_count.restype = C.c_int
    L.sfgpu_create.argtypes = [C.POINTER(vp), C.POINTER(Opt), vp]
    L.sfgpu_set_ref.argtypes = [vp, C.c_int32, vp, vp, vp, vp, vp]
    L.sfgpu_submit.argtypes = [vp, C.c_int32, C.c_int32, vp, vp, vp, vp, vp]
    L.sfgpu_submit_reads.argtypes = [vp, C.c_int32, C.c_int32, vp, vp, vp, vp, vp]
    L.sfgpu_submit_records.argtypes = [vp, C.c_int32, C.c_int32, vp, vp, C.c_int32, C.c_int32, vp, vp, vp, vp, vp, vp]
    L.sfgpu_slot_signal.argtypes = [vp, C.c_int32, C.c_int32, vp, C.c_int64]
    L.sfgpu_slot_signal.restype = C.c_int64
    L.sfgpu_resubmit.argtypes = [vp, C.c_int32]
    L.sfgpu_collect.argtypes = [vp, C.c_int32, vp]
    L.sfgpu_timing.argtypes = [vp, C.c_int32, C.POINTER(Timing)]
    L.sfgpu_destroy.argtypes = [vp]
    L.sfgpu_destroy.restype = None
    L.sfgpu_strerror.argtypes = [vp]
    L.sfgpu_strerror.restype = C.c_char_p
    L.sfgpu_ref_events.argtypes = [vp, C.c_int32, C.c_int32, vp, C.c_int32]
    L.sfgpu_event_table.argtypes = [vp, vp, C.c_int64, C.c_float, C.c_float, C.c_float, vp, vp, vp, C.c_int64]
    L.sfgpu_event_table.restype = C.c_int64
    L.sfgpu_query.argtypes = [vp, C.c_int32, C.c_int32, vp, C.c_int32]
    L.sfgpu_set_ref_events.argtypes = [vp, C.c_int32, C.c_int32, vp, vp]
    L.sfgpu_submit_queries.argtypes = [vp, C.c_int32, C.c_int32, vp, vp]
    L.sfgpu_collect_paths.argtypes = [vp, C.c_int32, vp, vp, vp, vp, vp]
    L.sfgpu_ref_columns.argtypes = [vp]
    L.sfgpu_ref_columns.restype = C.c_int64
    L.sfgpu_wave_reads.argtypes = [vp]
    L.sfgpu_wave_reads.restype = C.c_int32
    _lib = L
    return L


class SfgpuError(RuntimeError):
    pass


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """One GPU context: model + resident reference + batch slots (sfgpu_ctx)."""

    def __init__(self, level_mean: np.ndarray, kmer_size: int, flags: int = 0, query_size: int = 250,
                 prefix_size: int = 50, device: int = 0, n_slots: int = 2, ck_min_cols: int = 0,
                 min_window: int = 0, pore: int = 0, no_pairing: bool = False, warm_blocks: int = 0,
                 piece_periods: int = 0):
        L = lib()
        self._h = C.c_void_p()
        self.opt = Opt(device=device, flags=flags, query_size=query_size, prefix_size=prefix_size,
                       kmer_size=kmer_size, n_slots=n_slots, pore=pore)
        self.opt.reserved[0] = ck_min_cols  # test knob: checkpoint segments longer than this
        self.opt.reserved[1] = min_window   # test knob: restart distance of the start-coordinate pass
        self.opt.reserved[2] = warm_blocks  # test knob: warm-up of a piece of a split segment, in 64-column blocks
        # test knob: True = one read per warp even where pairing is the default, 2 = pair for every 128 < q <= 256
        self.opt.reserved[3] = int(no_pairing)
        self.opt.reserved[4] = piece_periods  # test knob: piece length in checkpoint periods (< 0: never split)
        lm = np.ascontiguousarray(level_mean, dtype=np.float32)
        assert lm.shape[0] == 4 ** kmer_size
        rc = L.sfgpu_create(C.byref(self._h), C.byref(self.opt), _ptr(lm))
        if rc != 0:
            raise SfgpuError(f"sfgpu_create failed ({rc}): {L.sfgpu_strerror(None).decode()}")
        self.flags = flags
        self.num_ref = 0
        self._n = {}

    def _check(self, rc, what):
        if rc < 0:
            raise SfgpuError(f"{what} failed ({rc}): {lib().sfgpu_strerror(self._h).decode()}")
        return rc

    def set_ref(self, seqs):
        seqs = [s if isinstance(s, bytes) else s.encode() for s in seqs]
        bases = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy()
        off = np.zeros(len(seqs) + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(s) for s in seqs])
        self.ref_lengths = np.zeros(len(seqs), dtype=np.int32)
        self.ref_seq_lengths = np.zeros(len(seqs), dtype=np.int32)
        self.ref_st_offset = np.zeros(len(seqs), dtype=np.int32)
        self._check(lib().sfgpu_set_ref(self._h, len(seqs), _ptr(bases), _ptr(off), _ptr(self.ref_lengths),
                                        _ptr(self.ref_seq_lengths), _ptr(self.ref_st_offset)), "sfgpu_set_ref")
        self.num_ref = len(seqs)

    def set_ref_events(self, fwd, rev=None):
        """caller-made event arrays (lists of float32 arrays); rev=None: single strand"""
        parts, off = [], [0]
        for i, f in enumerate(fwd):
            parts.append(np.asarray(f, dtype=np.float32))
            off.append(off[-1] + len(f))
            if rev is not None:
                assert len(rev[i]) == len(f)
                parts.append(np.asarray(rev[i], dtype=np.float32))
                off.append(off[-1] + len(f))
        ev = np.ascontiguousarray(np.concatenate(parts))
        offs = np.array(off, dtype=np.int64)
        self._check(lib().sfgpu_set_ref_events(self._h, len(fwd), int(rev is not None), _ptr(ev), _ptr(offs)),
                    "sfgpu_set_ref_events")
        self.num_ref = len(fwd)
        self.ref_lengths = np.array([len(f) for f in fwd], dtype=np.int32)

    def submit_queries(self, slot, queries):
        """list of float32 query arrays (each <= query_size long)"""
        q = self.opt.query_size
        n = len(queries)
        buf = np.zeros((max(n, 1), q), dtype=np.float32)
        ql = np.zeros(max(n, 1), dtype=np.int32)
        for i, x in enumerate(queries):
            buf[i, :len(x)] = x
            ql[i] = len(x)
        self._check(lib().sfgpu_submit_queries(self._h, slot, n, _ptr(buf), _ptr(ql)), "sfgpu_submit_queries")
        self._n[slot] = n

    def align_queries(self, queries, slot: int = 0) -> np.ndarray:
        self.submit_queries(slot, queries)
        return self.collect(slot)

    @staticmethod
    def pack(signals, scalings):
        """list of int16 arrays + list of scaling dicts -> the flat arrays sfgpu_submit takes"""
        n = len(signals)
        off = np.zeros(n + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(s) for s in signals])
        flat = np.concatenate([np.asarray(s, dtype=np.int16) for s in signals]) if n else np.zeros(0, np.int16)
        dig = np.array([sc["digitisation"] for sc in scalings], dtype=np.float32)
        offs = np.array([sc["offset"] for sc in scalings], dtype=np.float32)
        rng = np.array([sc["range"] for sc in scalings], dtype=np.float32)
        return np.ascontiguousarray(flat), off, dig, offs, rng

    def submit(self, slot, flat, off, dig, offs, rng):
        n = off.shape[0] - 1
        self._check(lib().sfgpu_submit(self._h, slot, n, _ptr(flat), _ptr(off), _ptr(dig), _ptr(offs), _ptr(rng)),
                    "sfgpu_submit")
        self._n[slot] = n

    def submit_reads(self, slot, signals, scalings):
        """one buffer per read (sfgpu_submit_reads): no gather copy on the caller's side"""
        n = len(signals)
        keep = [np.ascontiguousarray(s, dtype=np.int16) for s in signals]
        ptrs = (C.c_void_p * max(n, 1))(*[k.ctypes.data for k in keep])
        lens = np.array([len(k) for k in keep] or [0], dtype=np.int64)
        dig = np.array([sc["digitisation"] for sc in scalings] or [0], dtype=np.float32)
        offs = np.array([sc["offset"] for sc in scalings] or [0], dtype=np.float32)
        rng = np.array([sc["range"] for sc in scalings] or [0], dtype=np.float32)
        self._check(lib().sfgpu_submit_reads(self._h, slot, n, ptrs, _ptr(lens), _ptr(dig), _ptr(offs), _ptr(rng)),
                    "sfgpu_submit_reads")
        self._n[slot] = n

    def submit_records(self, slot, records, record_press, signal_press, sig_pos, sig_bytes, n_samples, scalings):
        """BLOW5 records as they lie in the file (list of bytes objects): inflate + signal decoding on the device"""
        n = len(records)
        keep = [np.frombuffer(r, dtype=np.uint8) if len(r) else np.zeros(0, np.uint8) for r in records]
        ptrs = (C.c_void_p * max(n, 1))(*[(k.ctypes.data if k.size else 0) for k in keep])
        nbytes = np.array([k.size for k in keep] or [0], dtype=np.int64)
        spos = np.array(list(sig_pos) or [0], dtype=np.int32)
        sbytes = np.array(list(sig_bytes) or [0], dtype=np.int64)
        ns = np.array(list(n_samples) or [0], dtype=np.int64)
        dig = np.array([sc["digitisation"] for sc in scalings] or [0], dtype=np.float32)
        offs = np.array([sc["offset"] for sc in scalings] or [0], dtype=np.float32)
        rng = np.array([sc["range"] for sc in scalings] or [0], dtype=np.float32)
        self._check(lib().sfgpu_submit_records(self._h, slot, n, ptrs, _ptr(nbytes), int(record_press), int(signal_press),
                                               _ptr(spos), _ptr(sbytes), _ptr(ns), _ptr(dig), _ptr(offs), _ptr(rng)),
                    "sfgpu_submit_records")
        self._n[slot] = n

    def slot_signal(self, slot, read, n_samples) -> np.ndarray:
        out = np.zeros(max(int(n_samples), 1), dtype=np.int16)
        n = self._check(lib().sfgpu_slot_signal(self._h, slot, read, _ptr(out), out.shape[0]), "sfgpu_slot_signal")
        return out[:n]

    def resubmit(self, slot):
        self._check(lib().sfgpu_resubmit(self._h, slot), "sfgpu_resubmit")

    def collect(self, slot) -> np.ndarray:
        n = self._n[slot]
        out = np.zeros(max(n, 1), dtype=RESULT_DTYPE)
        self._check(lib().sfgpu_collect(self._h, slot, _ptr(out)), "sfgpu_collect")
        return out[:n]

    def collect_paths(self, slot, results: np.ndarray):
        """--sam: per read (px, py) of the winner's warping path in forward order plus the window's event
        starts / lengths; `results` is what collect(slot) returned"""
        n = len(results)
        q = self.opt.query_size
        need = np.where(results["qlen"] > 0, results["qlen"] + results["pos_end"] - results["pos_st"], 0).astype(np.int64)
        need = np.maximum(need, 0)
        off = np.zeros(n + 1, dtype=np.int64)
        off[1:] = np.cumsum(need)
        moves = np.zeros(max(int(off[-1]), 1), dtype=np.uint8)
        n_moves = np.zeros(max(n, 1), dtype=np.int32)
        ev_start = np.zeros(max(n, 1) * q, dtype=np.uint64)
        ev_len = np.zeros(max(n, 1) * q, dtype=np.float32)
        self._check(lib().sfgpu_collect_paths(self._h, slot, _ptr(off), _ptr(moves), _ptr(n_moves), _ptr(ev_start),
                                              _ptr(ev_len)), "sfgpu_collect_paths")
        paths = []
        for i in range(n):
            if n_moves[i] < 0:
                paths.append(None)
                continue
            mv = moves[off[i]:off[i] + n_moves[i]]
            di = np.where(mv == 1, 0, 1)
            dj = np.where(mv == 2, 0, 1)
            px = int(results["qlen"][i]) - 1 - np.concatenate([[0], np.cumsum(di)])
            py = int(results["pos_end"][i]) - np.concatenate([[0], np.cumsum(dj)])
            paths.append((px[::-1].astype(np.int32), py[::-1].astype(np.int32)))
        return paths, ev_start.reshape(-1, q)[:n], ev_len.reshape(-1, q)[:n]

    def timing(self, slot) -> Timing:
        t = Timing()
        self._check(lib().sfgpu_timing(self._h, slot, C.byref(t)), "sfgpu_timing")
        return t

    def map_batch(self, signals, scalings, slot: int = 0) -> np.ndarray:
        self.submit(slot, *self.pack(signals, scalings))
        return self.collect(slot)

    def ref_events(self, rid: int, strand: int) -> np.ndarray:
        out = np.zeros(int(self.ref_lengths[rid]), dtype=np.float32)
        n = self._check(lib().sfgpu_ref_events(self._h, rid, strand, _ptr(out), out.shape[0]), "sfgpu_ref_events")
        return out[:n]

    def event_table(self, signal, scaling):
        sig = np.ascontiguousarray(signal, dtype=np.int16)
        cap = sig.shape[0] + 1
        start = np.zeros(cap, dtype=np.uint64)
        length = np.zeros(cap, dtype=np.float32)
        mean = np.zeros(cap, dtype=np.float32)
        n = self._check(lib().sfgpu_event_table(self._h, _ptr(sig), sig.shape[0], np.float32(scaling["digitisation"]),
                                                np.float32(scaling["offset"]), np.float32(scaling["range"]),
                                                _ptr(start), _ptr(length), _ptr(mean), cap), "sfgpu_event_table")
        return start[:n], length[:n], mean[:n]

    def query(self, slot: int, read: int) -> np.ndarray:
        out = np.zeros(self.opt.query_size, dtype=np.float32)
        n = self._check(lib().sfgpu_query(self._h, slot, read, _ptr(out), out.shape[0]), "sfgpu_query")
        return out[:n]

    @property
    def wave_reads(self) -> int:
        return lib().sfgpu_wave_reads(self._h)

    @property
    def ref_columns(self) -> int:
        return lib().sfgpu_ref_columns(self._h)

    def close(self):
        if self._h:
            lib().sfgpu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
