"""Deterministic synthetic inputs for the `dtw` path.

The real ONT pore-model tables (reference src/model.h) are absent from the
reference mount, so every run -- reference binary, oracle and GPU path -- is
driven by a seeded synthetic k-mer model written in the `--kmer-model` text
format (reference src/model.c:60-109: optional `#k\\t<K>` line, header line,
then 4^K positional rows `kmer\\tlevel_mean\\tlevel_stdv...`).

All model values are multiples of 1/64 so that the `%.6f` text form parses back
to exactly the same fp32 value on every path (no decimal->binary double rounding).
"""
from __future__ import annotations

import os
import numpy as np

_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)

DNA_SCALING = dict(digitisation=8192.0, range=1402.882, offset=10.0, sampling_rate=4000.0)
RNA_SCALING = dict(digitisation=2048.0, range=548.788, offset=-240.0, sampling_rate=3000.0)


def make_model(k: int, seed: int = 7):
    """level_mean in [60,130], level_stdv in [1,3], both on a 1/64 grid."""
    rng = np.random.default_rng(seed * 1000003 + k)
    n = 4 ** k
    mean = (60.0 + rng.integers(0, 70 * 64, size=n) / 64.0).astype(np.float32)
    stdv = (1.0 + rng.integers(0, 2 * 64, size=n) / 64.0).astype(np.float32)
    return mean, stdv


def write_model_file(path: str, k: int, mean: np.ndarray, stdv: np.ndarray) -> None:
    n = 4 ** k
    assert mean.shape[0] == n
    idx = np.arange(n)
    # k-mer text is ignored by the reader (rows are positional) but keep it right
    cols = [(idx >> (2 * (k - 1 - i))) & 3 for i in range(k)]
    kmers = _BASES[np.stack(cols, axis=1)].view(f"S{k}").ravel()
    with open(path, "w") as f:
        f.write(f"#k\t{k}\n")
        f.write("kmer\tlevel_mean\tlevel_stdv\tsd_mean\tsd_stdv\n")
        lines = [f"{kmers[i].decode()}\t{mean[i]:.6f}\t{stdv[i]:.6f}\t0.000000\t0.000000\n" for i in range(n)]
        f.write("".join(lines))


def random_sequence(n: int, rng: np.random.Generator) -> bytes:
    return _BASES[rng.integers(0, 4, size=n)].tobytes()


def transcriptome(n: int, seed: int, lo: int = 400, hi: int = 4000):
    """n seeded random transcripts with lengths uniform in [lo, hi] (BASELINE.json configs[4]: many short
    references); returns (names, sequences)"""
    rng = np.random.default_rng(seed)
    lens = rng.integers(lo, hi + 1, size=n)
    flat = _BASES[rng.integers(0, 4, size=int(lens.sum()))]
    offs = np.zeros(n + 1, dtype=np.int64)
    offs[1:] = np.cumsum(lens)
    return [f"tx{i:06d}" for i in range(n)], [flat[offs[i]:offs[i + 1]].tobytes() for i in range(n)]


def write_fasta(path: str, names, seqs, width: int = 80) -> None:
    with open(path, "w") as f:
        for name, s in zip(names, seqs):
            if isinstance(s, bytes):
                s = s.decode()
            f.write(f">{name}\n")
            for i in range(0, len(s), width):
                f.write(s[i:i + width] + "\n")


def kmer_ranks(seq: bytes, k: int) -> np.ndarray:
    """rank of every k-mer of seq (A,C,G,T -> 0..3, anything else -> 0), cf. ref.h:13-41"""
    lut = np.zeros(256, dtype=np.int64)
    for ch, v in ((b"A", 0), (b"C", 1), (b"G", 2), (b"T", 3), (b"a", 0), (b"c", 1), (b"g", 2), (b"t", 3)):
        lut[ch[0]] = v
    b = lut[np.frombuffer(seq, dtype=np.uint8)]
    n = len(seq) + 1 - k
    r = np.zeros(n, dtype=np.int64)
    for i in range(k):
        r = (r << 2) | b[i:i + n]
    return r


def revcomp(seq: bytes) -> bytes:
    tbl = bytearray(b"T" * 256)  # ref.h:62-65: anything unknown -> 'T'
    for a, b in ((b"A", b"T"), (b"C", b"G"), (b"G", b"C"), (b"T", b"A"), (b"a", b"T"), (b"c", b"G"), (b"g", b"C"), (b"t", b"A")):
        tbl[a[0]] = b[0]
    return seq.translate(bytes(tbl))[::-1]


def simulate_read(levels: np.ndarray, rng: np.random.Generator, scaling: dict,
                  noise_pa: float = 1.5, mean_extra_dwell: float = 8.0, min_dwell: int = 2) -> np.ndarray:
    """Squiggle for a sequence of k-mer levels: dwell = min_dwell + Geom, N(0,noise) pA noise,
    quantised to int16 ADC counts with the given scaling (pa = (raw + offset) * range / digitisation)."""
    dwell = min_dwell + rng.geometric(1.0 / (mean_extra_dwell + 1.0), size=levels.shape[0]) - 1
    pa = np.repeat(levels.astype(np.float64), dwell)
    pa = pa + rng.normal(0.0, noise_pa, size=pa.shape[0])
    raw = np.rint(pa * scaling["digitisation"] / scaling["range"] - scaling["offset"])
    return np.clip(raw, -32768, 32767).astype(np.int16)


def simulate_reads(seqs, k: int, level_mean: np.ndarray, n_reads: int, seed: int, rna: bool = False,
                   bases_per_read: int = 450, both_strands: bool = True, scaling: dict | None = None,
                   adaptor_events: int = 0, min_samples: int = 0):
    """Draw n_reads loci from the given sequences and simulate their raw signal.

    DNA: a locus on either strand, read 5'->3' of that strand.
    RNA: the 3' end of a transcript, sequenced 3'->5' (so k-mer levels are emitted in reverse).
    Returns (list of int16 arrays, list of (contig index, strand, start) truth tuples)."""
    rng = np.random.default_rng(seed)
    scaling = scaling or (RNA_SCALING if rna else DNA_SCALING)
    fwd_ranks = [kmer_ranks(s, k) for s in seqs]
    rev_ranks = None if rna else [kmer_ranks(revcomp(s), k) for s in seqs]
    sigs, truth = [], []
    for _ in range(n_reads):
        ci = int(rng.integers(0, len(seqs)))
        n_k = fwd_ranks[ci].shape[0]
        span = min(bases_per_read, n_k)
        if rna:
            st = n_k - span
            lv = level_mean[fwd_ranks[ci][st:st + span]][::-1]
            strand = "+"
        else:
            st = int(rng.integers(0, n_k - span + 1))
            if both_strands and rng.integers(0, 2):
                lv = level_mean[rev_ranks[ci][st:st + span]]
                strand = "-"
            else:
                lv = level_mean[fwd_ranks[ci][st:st + span]]
                strand = "+"
        if adaptor_events:
            lv = np.concatenate([rng.uniform(70, 110, size=adaptor_events).astype(np.float32), lv])
        min_dwell = 6 if rna else 2
        if min_samples:  # the reference's (dead) MAD trim has UB on reads shorter than ~500 samples
            min_dwell = max(min_dwell, -(-min_samples // max(1, lv.shape[0])))
        sigs.append(simulate_read(lv, rng, scaling, mean_extra_dwell=(18.0 if rna else 8.0), min_dwell=min_dwell))
        truth.append((ci, strand, st))
    return sigs, truth


def write_slow5_ascii(path: str, read_ids, signals, rna: bool = False, kit: str | None = None,
                      scaling: dict | None = None, scalings=None) -> None:
    """ASCII SLOW5 v0.2.0 that the reference's slow5_open accepts."""
    scaling = scaling or (RNA_SCALING if rna else DNA_SCALING)
    kit = kit or ("sqk-rna002" if rna else "sqk-lsk109")
    with open(path, "w") as f:
        f.write("#slow5_version\t0.2.0\n#num_read_groups\t1\n")
        f.write(f"@experiment_type\t{'rna' if rna else 'genomic_dna'}\n")
        f.write(f"@sequencing_kit\t{kit}\n")
        f.write("#char*\tuint32_t\tdouble\tdouble\tdouble\tdouble\tuint64_t\tint16_t*\n")
        f.write("#read_id\tread_group\tdigitisation\toffset\trange\tsampling_rate\tlen_raw_signal\traw_signal\n")
        for i, (rid, sig) in enumerate(zip(read_ids, signals)):
            sc = scalings[i] if scalings is not None else scaling
            f.write(f"{rid}\t0\t{_g(sc['digitisation'])}\t{_g(sc['offset'])}\t{_g(sc['range'])}\t"
                    f"{_g(sc['sampling_rate'])}\t{len(sig)}\t")
            f.write(",".join(map(str, sig.tolist())))
            f.write("\n")


def _g(x: float) -> str:
    # shortest decimal that round-trips the double (slow5 parses with strtod)
    return repr(float(x))


def tmpdir() -> str:
    d = os.environ.get("SIGFISH_B200_TMP", "/tmp/sigfish_b200")
    os.makedirs(d, exist_ok=True)
    return d


# ---------------------------------------------------------------- BLOW5 writer (test inputs)

def _svb_zd_encode(sig: np.ndarray) -> bytes:
    """StreamVByte(zigzag(delta)) of an int16 signal, the BLOW5 'svb-zd' signal codec:
    u32 count, ceil(count/4) key bytes (2 bits per value: byte length - 1), then the data bytes."""
    x = sig.astype(np.int64)
    d = np.diff(x, prepend=0)
    z = ((d << 1) ^ (d >> 63)).astype(np.uint32)
    nbytes = np.where(z < (1 << 8), 1, np.where(z < (1 << 16), 2, np.where(z < (1 << 24), 3, 4))).astype(np.uint8)
    n = z.shape[0]
    codes = np.zeros(((n + 3) // 4) * 4, dtype=np.uint8)
    codes[:n] = nbytes - 1
    c4 = codes.reshape(-1, 4)
    keys = (c4[:, 0] | (c4[:, 1] << 2) | (c4[:, 2] << 4) | (c4[:, 3] << 6)).astype(np.uint8)
    raw = z.view(np.uint8).reshape(-1, 4)  # little endian
    mask = np.arange(4)[None, :] < nbytes[:, None]
    data = raw[mask]
    return np.uint32(n).tobytes() + keys.tobytes() + data.tobytes()


def write_blow5(path: str, read_ids, signals, rna: bool = False, kit: str | None = None,
                scaling: dict | None = None, scalings=None, record_zlib: bool = True, signal_svb: bool = True) -> None:
    """BLOW5 v0.2.0 file (zlib records, svb-zd signal by default) with the primary fields only."""
    import struct
    import zlib
    scaling = scaling or (RNA_SCALING if rna else DNA_SCALING)
    kit = kit or ("sqk-rna002" if rna else "sqk-lsk109")
    # the text part of a BLOW5 header holds only the @attributes and the two column lines: version, compression
    # and the number of read groups live in the binary part
    hdr = (f"@experiment_type\t{'rna' if rna else 'genomic_dna'}\n"
           f"@sequencing_kit\t{kit}\n#char*\tuint32_t\tdouble\tdouble\tdouble\tdouble\tuint64_t\tint16_t*\n"
           "#read_id\tread_group\tdigitisation\toffset\trange\tsampling_rate\tlen_raw_signal\traw_signal\n").encode()
    with open(path, "wb") as f:
        head = b"BLOW5\x01" + bytes([0, 2, 0]) + bytes([1 if record_zlib else 0]) + struct.pack("<I", 1) + \
            bytes([1 if signal_svb else 0])
        f.write(head + b"\x00" * (64 - len(head)))
        f.write(struct.pack("<I", len(hdr)) + hdr)
        for i, (rid, sig) in enumerate(zip(read_ids, signals)):
            sc = scalings[i] if scalings is not None else scaling
            sig = np.ascontiguousarray(sig, dtype=np.int16)
            body = sig.tobytes() if not signal_svb else _svb_zd_encode(sig)
            n_field = len(sig) if not signal_svb else len(body)
            rid_b = rid.encode()
            rec = struct.pack("<H", len(rid_b)) + rid_b + struct.pack("<I", 0) + \
                struct.pack("<dddd", sc["digitisation"], sc["offset"], sc["range"], sc["sampling_rate"]) + \
                struct.pack("<Q", n_field) + body
            if record_zlib:
                rec = zlib.compress(rec)
            f.write(struct.pack("<Q", len(rec)) + rec)
        f.write(b"5WOLB")


def simulate_rna_reads_with_tail(seqs, k: int, level_mean: np.ndarray, n_reads: int, seed: int,
                                 bases_per_read: int = 700, scaling: dict | None = None):
    """direct-RNA-like reads for the automatic query start (-p -1): a low-current adaptor stretch, a flat
    poly-A plateau about 30 pA above it, then the transcript body read 3'->5'.  Every fourth read has no
    tail at all (the detector must fail and fall back to 50 events)."""
    rng = np.random.default_rng(seed)
    scaling = scaling or RNA_SCALING
    ranks_ = [kmer_ranks(s, k) for s in seqs]
    sigs, truth = [], []
    for r in range(n_reads):
        ci = int(rng.integers(0, len(seqs)))
        n_k = ranks_[ci].shape[0]
        span = min(bases_per_read, n_k)
        lv = level_mean[ranks_[ci][n_k - span:]][::-1].astype(np.float64)
        dwell = 6 + rng.geometric(1.0 / 19.0, size=lv.shape[0]) - 1
        body = np.repeat(lv, dwell) + rng.normal(0.0, 1.5, size=int(dwell.sum()))
        if r % 4 != 3:
            a_len = int(rng.integers(2600, 6000))
            p_len = int(rng.integers(500, 1800))
            a_lvl = float(rng.uniform(62.0, 72.0))
            adaptor = rng.normal(a_lvl, 2.5, size=a_len)
            polya = rng.normal(a_lvl + 30.0 + rng.uniform(-6, 6), 1.5, size=p_len)
            pa = np.concatenate([adaptor, polya, body])
        else:
            pa = body
        raw = np.rint(pa * scaling["digitisation"] / scaling["range"] - scaling["offset"])
        sigs.append(np.clip(raw, -32768, 32767).astype(np.int16))
        truth.append((ci, "+", n_k - span))
    return sigs, truth
