"""Test-side bindings: the CPU oracle (oracle/liboracle.so), and -- when it was built in the
container that has /root/reference -- the unmodified reference (oracle/_ref/).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module.
"""
from __future__ import annotations

import ctypes as C
import gzip
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "sigfish")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libsigfish_ref.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")

F_RNA, F_DTW, F_INV, F_REF, F_END = 0x001, 0x002, 0x004, 0x010, 0x020

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i16p = np.ctypeslib.ndpointer(dtype=np.int16, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


class OrcEvent(C.Structure):
    _fields_ = [("start", C.c_uint64), ("length", C.c_float), ("mean", C.c_float), ("stdv", C.c_float)]


EVENT_DTYPE = np.dtype([("start", "<u8"), ("length", "<f4"), ("mean", "<f4"), ("stdv", "<f4")], align=True)


class OrcHit(C.Structure):
    _fields_ = [("mapped", C.c_int32), ("status", C.c_int32), ("n_events", C.c_int64),
                ("qstart", C.c_int64), ("qend", C.c_int64), ("start_raw", C.c_uint64),
                ("end_raw", C.c_uint64), ("rid", C.c_int32), ("pos_st", C.c_int32),
                ("pos_end", C.c_int32), ("raw_pos_st", C.c_int32), ("raw_pos_end", C.c_int32),
                ("score", C.c_float), ("score2", C.c_float), ("mapq", C.c_int32), ("strand", C.c_char)]


def build_oracle() -> None:
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "oracle"], check=True)


_oracle = None


def oracle():
    global _oracle
    if _oracle is not None:
        return _oracle
    if not os.path.exists(ORACLE_SO):
        build_oracle()
    L = C.CDLL(ORACLE_SO)
    L.orc_to_picoamps.argtypes = [_i16p, C.c_int64, C.c_float, C.c_float, C.c_float, _f32p]
    L.orc_prefix_sums.argtypes = [_f32p, C.c_int64, _f64p, _f64p]
    L.orc_tstat.argtypes = [_f64p, _f64p, C.c_int64, C.c_int64, _f32p]
    L.orc_peaks.argtypes = [_f32p, _f32p, C.c_int64, C.c_int, _u64p]
    L.orc_peaks.restype = C.c_int64
    L.orc_detect_events.argtypes = [_i16p, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_int,
                                    C.POINTER(C.POINTER(OrcEvent))]
    L.orc_detect_events.restype = C.c_int64
    L.orc_ref_build.argtypes = [C.c_int32, C.POINTER(C.c_char_p), _i32p, _f32p, C.c_int32, C.c_uint32, C.c_int32]
    L.orc_ref_build.restype = C.c_void_p
    L.orc_ref_free.argtypes = [C.c_void_p]
    L.orc_ref_from_events.argtypes = [C.c_int32, C.c_int32, _i32p, _f32p, C.c_void_p]
    L.orc_ref_from_events.restype = C.c_void_p
    L.orc_align_means.argtypes = [C.c_void_p, _f32p, C.c_int32, C.c_uint32, C.POINTER(OrcHit)]
    for fn in ("orc_ref_len", "orc_ref_offset"):
        getattr(L, fn).argtypes = [C.c_void_p, C.c_int32]
        getattr(L, fn).restype = C.c_int32
    for fn in ("orc_ref_fwd", "orc_ref_rev"):
        getattr(L, fn).argtypes = [C.c_void_p, C.c_int32]
        getattr(L, fn).restype = C.POINTER(C.c_float)
    L.orc_subsequence.argtypes = [_f32p, _f32p, C.c_int, C.c_int, _f32p]
    L.orc_std_dtw.argtypes = [_f32p, _f32p, C.c_int, C.c_int, _f32p]
    L.orc_std_dtw.restype = C.c_float
    L.orc_path_start.argtypes = [_f32p, C.c_int, C.c_int, C.c_int]
    L.orc_path_start.restype = C.c_int32
    L.orc_path_full.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, _i32p, _i32p]
    L.orc_path_full.restype = C.c_int32
    L.orc_window_normalise.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_uint32, C.c_int32, C.c_int32,
                                       C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
    L.orc_window_normalise.restype = C.c_int
    L.orc_map_read.argtypes = [C.c_void_p, _i16p, C.c_int64, C.c_float, C.c_float, C.c_float,
                               C.c_uint32, C.c_int32, C.c_int32, C.POINTER(OrcHit)]
    L.orc_paf_line.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(OrcHit), C.c_char_p, C.c_char_p,
                               C.c_int32, C.c_int64]
    L.orc_paf_line.restype = C.c_int
    L.orc_free.argtypes = [C.c_void_p]
    L.orc_map_read_sam.argtypes = [C.c_void_p, _i16p, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_uint32, C.c_int32,
                                   C.c_int32, C.c_char_p, C.POINTER(C.c_char_p), C.c_char_p, C.c_size_t]
    L.orc_polya_end_sample.argtypes = [_i16p, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_int]
    L.orc_polya_end_sample.restype = C.c_int64
    _oracle = L
    return L


# ---------------------------------------------------------------- oracle wrappers

def orc_events(raw: np.ndarray, dig: float, off: float, rng: float, rna: bool) -> np.ndarray:
    """structured array of events for one read (oracle)"""
    L = oracle()
    raw = np.ascontiguousarray(raw, dtype=np.int16)
    out = C.POINTER(OrcEvent)()
    n = L.orc_detect_events(raw, raw.shape[0], np.float32(dig), np.float32(off), np.float32(rng), int(rna), C.byref(out))
    if n <= 0:
        return np.zeros(0, dtype=EVENT_DTYPE)
    buf = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_uint8)), shape=(n * C.sizeof(OrcEvent),))
    ev = buf.view(EVENT_DTYPE).copy()
    L.orc_free(out)
    return ev


class OracleRef:
    def __init__(self, seqs, level_mean: np.ndarray, k: int, flags: int, q: int):
        L = oracle()
        self.seqs = [s if isinstance(s, bytes) else s.encode() for s in seqs]
        arr = (C.c_char_p * len(self.seqs))(*self.seqs)
        lens = np.array([len(s) for s in self.seqs], dtype=np.int32)
        self.level_mean = np.ascontiguousarray(level_mean, dtype=np.float32)
        self.h = L.orc_ref_build(len(self.seqs), arr, lens, self.level_mean, k, flags, q)
        self.n = len(self.seqs)
        self.flags = flags
        self.seq_lens = lens

    def length(self, i):
        return oracle().orc_ref_len(self.h, i)

    def offset(self, i):
        return oracle().orc_ref_offset(self.h, i)

    def fwd(self, i):
        return np.ctypeslib.as_array(oracle().orc_ref_fwd(self.h, i), shape=(self.length(i),)).copy()

    def rev(self, i):
        return np.ctypeslib.as_array(oracle().orc_ref_rev(self.h, i), shape=(self.length(i),)).copy()

    def close(self):
        if self.h:
            oracle().orc_ref_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class OracleEventRef:
    """oracle reference made of caller-given event arrays (lists of float32 arrays)"""

    def __init__(self, fwd, rev=None):
        lens = np.array([len(f) for f in fwd], dtype=np.int32)
        ff = np.ascontiguousarray(np.concatenate([np.asarray(f, np.float32) for f in fwd]))
        rr = None
        if rev is not None:
            rr = np.ascontiguousarray(np.concatenate([np.asarray(f, np.float32) for f in rev]))
        self.h = oracle().orc_ref_from_events(len(fwd), int(rev is not None), lens, ff,
                                              rr.ctypes.data_as(C.c_void_p) if rr is not None else None)

    def align(self, means, flags) -> OrcHit:
        hit = OrcHit()
        m = np.ascontiguousarray(means, dtype=np.float32)
        oracle().orc_align_means(self.h, m, m.shape[0], flags, C.byref(hit))
        return hit

    def close(self):
        if self.h:
            oracle().orc_ref_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def orc_map(ref: OracleRef, raw: np.ndarray, dig: float, off: float, rng: float, flags: int, q: int, p: int) -> OrcHit:
    hit = OrcHit()
    raw = np.ascontiguousarray(raw, dtype=np.int16)
    oracle().orc_map_read(ref.h, raw, raw.shape[0], np.float32(dig), np.float32(off), np.float32(rng), flags, q, p, C.byref(hit))
    return hit


def orc_paf(hit: OrcHit, read_id: str, rname: str, ref_seq_len: int, len_raw: int) -> str:
    buf = C.create_string_buffer(4096)
    n = oracle().orc_paf_line(buf, 4096, C.byref(hit), read_id.encode(), rname.encode(), ref_seq_len, len_raw)
    return buf.raw[:n].decode()


def oracle_paf(names, seqs, level_mean, k, read_ids, signals, scalings, flags, q=250, p=50) -> str:
    """Whole `sigfish dtw` run through the oracle -> PAF text."""
    ref = OracleRef(seqs, level_mean, k, flags, q)
    out = []
    for rid, sig, sc in zip(read_ids, signals, scalings):
        hit = orc_map(ref, sig, sc["digitisation"], sc["offset"], sc["range"], flags, q, p)
        if hit.mapped:
            out.append(orc_paf(hit, rid, names[hit.rid], int(ref.seq_lens[hit.rid]), len(sig)))
    ref.close()
    return "".join(out)


def oracle_sam(names, seqs, level_mean, k, read_ids, signals, scalings, flags, q=250, p=50) -> str:
    """Whole `sigfish dtw --sam` run through the oracle -> SAM text (header + records)."""
    ref = OracleRef(seqs, level_mean, k, flags, q)
    out = [f"@SQ\tSN:{names[i]}\tLN:{ref.length(i)}\n" for i in range(ref.n)]
    arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
    buf = C.create_string_buffer(1 << 20)
    for rid, sig, sc in zip(read_ids, signals, scalings):
        raw = np.ascontiguousarray(sig, dtype=np.int16)
        n = oracle().orc_map_read_sam(ref.h, raw, raw.shape[0], np.float32(sc["digitisation"]), np.float32(sc["offset"]),
                                      np.float32(sc["range"]), flags, q, p, rid.encode(), arr, buf, 1 << 20)
        out.append(buf.raw[:n].decode())
    ref.close()
    return "".join(out)


# ---------------------------------------------------------------- reference binary

def have_ref_bin() -> bool:
    return os.path.exists(REF_BIN)


def flags_to_cli(flags: int) -> list[str]:
    a = []
    if flags & F_RNA:
        a.append("--rna")
    if flags & F_DTW:
        a.append("--dtw-std")
    if flags & F_INV:
        a.append("--invert")
    if flags & F_REF:
        a.append("--full-ref")
    if flags & F_END:
        a.append("--from-end")
    return a


def run_ref(fasta: str, slow5: str, model: str, flags: int = 0, q: int = 250, p: int = 50, threads: int = 4,
            extra=()) -> str:
    cmd = [REF_BIN, "dtw", fasta, slow5, "--kmer-model", model, "-t", str(threads), "-q", str(q), "-p", str(p)]
    cmd += flags_to_cli(flags) + list(extra)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"reference failed ({r.returncode}): {r.stderr[-2000:]}")
    return r.stdout


def read_fasta(path: str):
    op = gzip.open if path.endswith(".gz") else open
    names, seqs, cur = [], [], []
    with op(path, "rt") as f:
        for line in f:
            line = line.rstrip("\n").rstrip("\r")
            if line.startswith(">"):
                if names:
                    seqs.append("".join(cur).encode())
                names.append(line[1:].split()[0] if len(line) > 1 else "")
                cur = []
            elif names:
                cur.append(line)
    if names:
        seqs.append("".join(cur).encode())
    return names, seqs


# fasta fixtures that are generated from a seed instead of being stored: name -> (transcripts, seed, lo, hi)
GENERATED_FASTA = {"gen_rna004_tx2000": (2000, 20240, 400, 4000)}


def case_fasta(c_or_name):
    """(names, sequences) of a golden case's reference: a committed .fa.gz, or a seeded transcriptome whose
    checksum make_golden.py recorded in cases.json"""
    name = c_or_name if isinstance(c_or_name, str) else c_or_name["fasta"]
    if name in GENERATED_FASTA:
        import hashlib
        from sigfish_b200 import synth
        n, seed, lo, hi = GENERATED_FASTA[name]
        names, seqs = synth.transcriptome(n, seed, lo, hi)
        if not isinstance(c_or_name, str) and "fasta_sha1" in c_or_name:
            got = hashlib.sha1(b"\n".join(seqs)).hexdigest()
            assert got == c_or_name["fasta_sha1"], "seeded transcriptome differs from the one the golden PAF was made with"
        return names, seqs
    return read_fasta(os.path.join(GOLDEN, name + ".fa.gz"))


def write_case_fasta(c_or_name, path: str) -> None:
    from sigfish_b200 import synth
    names, seqs = case_fasta(c_or_name)
    synth.write_fasta(path, names, seqs, width=60)


def case_reads(c):
    return load_reads_npz(os.path.join(GOLDEN, c["reads"] + ".npz"))


def case_kit(c):
    """sequencing_kit header value of the case's read file (selects the pore: src/sigfish.c:52-79)"""
    if c.get("kit"):
        return c["kit"]
    return "sqk-lsk114" if c["k"] == 9 else None


def case_oracle_flags(c) -> int:
    """oracle flag word: the option bits plus the oracle-only RNA004 bit (jnn parameter set)"""
    return c["flags"] | (0x400 if c.get("pore", 0) == 2 else 0)


def load_reads_npz(path: str):
    """fixture written by tests/golden/make_golden.py"""
    z = np.load(path, allow_pickle=False)
    ids = [s for s in z["read_ids"].tolist()]
    sig = z["signal"]
    offs = z["offsets"]
    sigs = [sig[offs[i]:offs[i + 1]] for i in range(len(ids))]
    sc = [dict(digitisation=float(z["digitisation"][i]), offset=float(z["offset"][i]),
               range=float(z["range"][i]), sampling_rate=float(z["sampling_rate"][i])) for i in range(len(ids))]
    return ids, sigs, sc
