#!/usr/bin/env python
"""Regenerate tests/golden/* from the UNMODIFIED reference.  Runs only in the build
container (needs /root/reference and oracle/_ref/ built by `make -C oracle ref`).

What it writes (all small, committed):
  * sp1_dna.npz, sequin_rna.npz  -- the raw int16 signals + scaling of the reference's bundled
    BLOW5 test reads (test/sp1_dna.blow5, test/sequin_rna.blow5), decoded with the reference's
    own slow5lib through ctypes.  Data, not code.
  * nCoV-2019.fa.gz, rnasequin.fa.gz -- the bundled FASTA references, gzip'd.
  * synth_*.npz -- seeded synthetic read sets (sigfish_b200.synth).
  * paf/<case>.paf -- stdout of `oracle/_ref/sigfish dtw ... --kmer-model <synthetic model>` for
    every case in CASES: the golden vectors that pin the oracle and the GPU path.
  * events_*.npz -- event tables from the reference's getevents() (events.c:557) called directly.

The synthetic k-mer models are regenerated from sigfish_b200.synth.make_model(k, seed=7).
"""
import ctypes as C
import gzip
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers as H  # noqa: E402
from sigfish_b200 import synth  # noqa: E402

REFDIR = "/root/reference/test"


class S5Rec(C.Structure):
    _fields_ = [("read_id_len", C.c_uint16), ("read_id", C.c_char_p), ("read_group", C.c_uint32),
                ("digitisation", C.c_double), ("offset", C.c_double), ("range", C.c_double),
                ("sampling_rate", C.c_double), ("len_raw_signal", C.c_uint64),
                ("raw_signal", C.POINTER(C.c_int16)), ("aux_map", C.c_void_p)]


class RefEvent(C.Structure):
    _fields_ = [("start", C.c_uint64), ("length", C.c_float), ("mean", C.c_float), ("stdv", C.c_float)]


class RefEventTable(C.Structure):
    _fields_ = [("n", C.c_size_t), ("start", C.c_size_t), ("end", C.c_size_t), ("event", C.POINTER(RefEvent))]


def ref_lib():
    L = C.CDLL(H.REF_SO)
    L.slow5_open.argtypes = [C.c_char_p, C.c_char_p]
    L.slow5_open.restype = C.c_void_p
    L.slow5_get_next.argtypes = [C.POINTER(C.POINTER(S5Rec)), C.c_void_p]
    L.slow5_rec_free.argtypes = [C.POINTER(S5Rec)]
    L.slow5_close.argtypes = [C.c_void_p]
    L.getevents.argtypes = [C.c_size_t, np.ctypeslib.ndpointer(dtype=np.float32), C.c_int8]
    L.getevents.restype = RefEventTable
    return L


def read_blow5(path):
    L = ref_lib()
    sp = L.slow5_open(path.encode(), b"r")
    assert sp, path
    rec = C.POINTER(S5Rec)()
    ids, sigs, sc = [], [], []
    while L.slow5_get_next(C.byref(rec), sp) >= 0:
        r = rec.contents
        ids.append(r.read_id.decode())
        sigs.append(np.ctypeslib.as_array(r.raw_signal, shape=(r.len_raw_signal,)).copy())
        sc.append(dict(digitisation=r.digitisation, offset=r.offset, range=r.range, sampling_rate=r.sampling_rate))
    L.slow5_rec_free(rec)
    L.slow5_close(sp)
    return ids, sigs, sc


def save_reads(path, ids, sigs, sc):
    offs = np.zeros(len(sigs) + 1, dtype=np.int64)
    offs[1:] = np.cumsum([len(s) for s in sigs])
    np.savez_compressed(path, read_ids=np.array(ids), signal=np.concatenate(sigs).astype(np.int16), offsets=offs,
                        digitisation=np.array([s["digitisation"] for s in sc]),
                        offset=np.array([s["offset"] for s in sc]), range=np.array([s["range"] for s in sc]),
                        sampling_rate=np.array([s["sampling_rate"] for s in sc]))


def ref_events(sig, sc, rna):
    """reference event_single (sigfish.c:330-353): pA conversion in numpy fp32, then getevents()"""
    L = ref_lib()
    unit = np.float32(sc["range"]) / np.float32(sc["digitisation"])
    pa = ((sig.astype(np.float32) + np.float32(sc["offset"])) * unit).astype(np.float32)
    et = L.getevents(len(pa), pa, 1 if rna else 0)
    out = np.zeros(et.n, dtype=H.EVENT_DTYPE)
    for i in range(et.n):
        e = et.event[i]
        out[i] = (e.start, e.length, e.mean, e.stdv)
    return out


# case name -> (reads fixture, fasta fixture, k, flags, q, p)
CASES = {
    "dna_sp1_default": ("sp1_dna", "nCoV-2019", 6, 0, 250, 50),
    "dna_sp1_from_end": ("sp1_dna", "nCoV-2019", 6, H.F_END, 250, 50),
    "dna_sp1_q100_p20": ("sp1_dna", "nCoV-2019", 6, 0, 100, 20),
    "dna_sp1_q300_p0": ("sp1_dna", "nCoV-2019", 6, 0, 300, 0),
    # query sizes that run in the two-reads-per-warp layout with 10 / 13 / 15 rows per lane
    "dna_synth48_q150": ("synth_dna48", "nCoV-2019", 6, 0, 150, 50),
    "dna_sp1_q200_p20": ("sp1_dna", "nCoV-2019", 6, 0, 200, 20),
    "dna_synth48_q230_from_end": ("synth_dna48", "nCoV-2019", 6, H.F_END, 230, 50),
    "rna_sequin_q180": ("sequin_rna", "rnasequin", 5, H.F_RNA, 180, 50),
    "dna_synth48": ("synth_dna48", "nCoV-2019", 6, 0, 250, 50),
    "dna_synth48_from_end": ("synth_dna48", "nCoV-2019", 6, H.F_END, 250, 50),
    "dna_multi_contig": ("synth_dna_multi", "synth_multi", 6, 0, 250, 50),
    "dna_short_reads": ("synth_dna_short", "nCoV-2019", 6, 0, 250, 50),
    "dna_short_reads_from_end": ("synth_dna_short", "nCoV-2019", 6, H.F_END, 250, 50),
    "dna_short_reads_p200": ("synth_dna_short", "nCoV-2019", 6, 0, 250, 200),
    "dna_short_reads_p200_from_end": ("synth_dna_short", "nCoV-2019", 6, H.F_END, 250, 200),
    "dna_r10_k9": ("synth_dna_k9", "synth_multi", 9, 0, 250, 50),
    "rna_sequin_default": ("sequin_rna", "rnasequin", 5, H.F_RNA, 250, 50),
    "rna_sequin_full_ref": ("sequin_rna", "rnasequin", 5, H.F_RNA | H.F_REF, 250, 50),
    "rna_sequin_dtw_std": ("sequin_rna", "rnasequin", 5, H.F_RNA | H.F_DTW, 250, 50),
    "rna_sequin_dtw_std_full_ref": ("sequin_rna", "rnasequin", 5, H.F_RNA | H.F_DTW | H.F_REF, 250, 50),
    "rna_sequin_invert": ("sequin_rna", "rnasequin", 5, H.F_RNA | H.F_INV, 250, 50),
    "rna_sequin_invert_full_ref": ("sequin_rna", "rnasequin", 5, H.F_RNA | H.F_INV | H.F_REF, 250, 50),
    "rna_sequin_from_end": ("sequin_rna", "rnasequin", 5, H.F_RNA | H.F_END, 250, 50),
    "rna_sequin_full_ref_from_end": ("sequin_rna", "rnasequin", 5, H.F_RNA | H.F_REF | H.F_END, 250, 50),
    "rna_sequin_q500": ("sequin_rna", "rnasequin", 5, H.F_RNA, 500, 50),
    "rna_synth32": ("synth_rna32", "rnasequin", 5, H.F_RNA, 250, 50),
    # automatic query start (jnn adaptor / poly-A finders); the first is the reference's own test
    # configuration (test/test.sh:70)
    "rna_sequin_q500_auto": ("sequin_rna", "rnasequin", 5, H.F_RNA, 500, -1),
    "rna_sequin_auto": ("sequin_rna", "rnasequin", 5, H.F_RNA, 250, -1),
    "rna_sequin_auto_full_ref": ("sequin_rna", "rnasequin", 5, H.F_RNA | H.F_REF, 250, -1),
    "rna_tail24_auto": ("synth_rna_tail24", "rnasequin", 5, H.F_RNA, 250, -1),
    "rna_tail24_auto_dtw_std": ("synth_rna_tail24", "rnasequin", 5, H.F_RNA | H.F_DTW, 250, -1),
    "rna_tail24_p50": ("synth_rna_tail24", "rnasequin", 5, H.F_RNA, 250, 50),
    # BASELINE.json configs[4] in small: RNA004 (9-mer RNA model, kit sqk-rna004 => pore auto-detected,
    # sigfish.c:62-64,123-135), a seeded transcriptome of 2000 transcripts of 400-4000 nt (not stored: regenerated
    # from its seed, checksum in cases.json), reads from the 3' ends
    "rna004_tx2000_invert": ("synth_rna004_24", "gen_rna004_tx2000", 9, H.F_RNA | H.F_INV, 250, 50),
    "rna004_tx2000_default": ("synth_rna004_24", "gen_rna004_tx2000", 9, H.F_RNA, 250, 50),
    "rna004_tx2000_invert_full_ref": ("synth_rna004_24", "gen_rna004_tx2000", 9, H.F_RNA | H.F_INV | H.F_REF, 250, 50),
    "rna004_tx2000_dtw_std": ("synth_rna004_24", "gen_rna004_tx2000", 9, H.F_RNA | H.F_DTW, 250, 50),
    "rna004_tail16_auto": ("synth_rna004_tail16", "gen_rna004_tx2000", 9, H.F_RNA, 250, -1),
}

# per-case extras: the sequencing_kit written into the read file's header and the pore it makes the reference
# detect (opt.pore_flag: selects the jnn parameter set of -p -1)
CASE_EXTRA = {c: dict(kit="sqk-rna004", pore=2) for c in CASES if c.startswith("rna004_")}
# 4.4 M reference columns per read: the single-threaded oracle checks only the first reads of this case (the
# GPU command line is compared with all golden lines)
CASE_EXTRA["rna004_tx2000_invert_full_ref"]["oracle_reads"] = 3
for _c in ("rna004_tx2000_default", "rna004_tx2000_dtw_std", "rna004_tail16_auto"):
    CASE_EXTRA[_c]["oracle_reads"] = 8  # keeps the CPU suite short; rna004_tx2000_invert runs all 24


# name -> (truth PAF, test PAF, options) for `sigfish eval`
EVAL_CASES = {
    "dna_synth48_vs_from_end": ("dna_synth48.paf", "dna_synth48_from_end.paf", []),
    "rna_default_vs_full_ref": ("rna_sequin_default.paf", "rna_sequin_full_ref.paf", []),
    "rna_default_vs_full_ref_tid_only": ("rna_sequin_default.paf", "rna_sequin_full_ref.paf", ["--tid-only"]),
    "disjoint_read_sets": ("dna_short_reads.paf", "dna_synth48.paf", []),
    "crafted": ("crafted_truth.paf", "crafted_test.paf", []),
    "crafted_no_secondary": ("crafted_truth.paf", "crafted_test.paf", ["--secondary", "no"]),
    "crafted_tid_only": ("crafted_truth.paf", "crafted_test.paf", ["--tid-only"]),
}

SAM_CASES = ["dna_sp1_default", "dna_sp1_from_end", "dna_synth48", "dna_multi_contig", "dna_short_reads", "dna_r10_k9",
             "rna_sequin_default", "rna_sequin_invert", "rna_sequin_full_ref", "rna_sequin_q500_auto",
             "rna_tail24_auto", "rna_synth32", "rna004_tx2000_invert", "rna004_tail16_auto"]


def only_cases(names):
    """adds the golden PAF of the named cases (fixtures must exist) without touching the other files"""
    tmp = synth.tmpdir()
    summary = json.load(open(os.path.join(HERE, "cases.json")))
    for case in names:
        reads, fasta, k, flags, q, p = CASES[case]
        mean, stdv = synth.make_model(k)
        synth.write_model_file(os.path.join(tmp, f"model_k{k}.txt"), k, mean, stdv)
        ids, sg, sc = H.load_reads_npz(os.path.join(HERE, reads + ".npz"))
        extra = CASE_EXTRA.get(case, {})
        s5 = os.path.join(tmp, reads + ("_" + extra["kit"] if extra else "") + ".slow5")
        synth.write_slow5_ascii(s5, ids, sg, rna=bool(flags & H.F_RNA), kit=extra.get("kit"), scalings=sc)
        fa = os.path.join(tmp, fasta + ".fa")
        with gzip.open(os.path.join(HERE, fasta + ".fa.gz"), "rb") as fi, open(fa, "wb") as fo:
            shutil.copyfileobj(fi, fo)
        paf = H.run_ref(fa, s5, os.path.join(tmp, f"model_k{k}.txt"), flags=flags, q=q, p=p)
        with open(os.path.join(HERE, "paf", case + ".paf"), "w") as f:
            f.write(paf)
        summary[case] = dict(reads=reads, fasta=fasta, k=k, flags=flags, q=q, p=p, rows=paf.count("\n"), **extra)
        print(case, summary[case]["rows"], "rows")
    with open(os.path.join(HERE, "cases.json"), "w") as f:
        json.dump(summary, f, indent=1, sort_keys=True)


def main():
    if "--only" in sys.argv:
        return only_cases(sys.argv[sys.argv.index("--only") + 1].split(","))
    os.makedirs(os.path.join(HERE, "paf"), exist_ok=True)
    tmp = synth.tmpdir()

    # 1. bundled data -> fixtures
    for name, src in (("sp1_dna", "sp1_dna.blow5"), ("sequin_rna", "sequin_rna.blow5")):
        ids, sigs, sc = read_blow5(os.path.join(REFDIR, src))
        save_reads(os.path.join(HERE, name + ".npz"), ids, sigs, sc)
    # the small DNA file itself (22 KB: zlib records, svb-zd signals, auxiliary fields, written by slow5lib): input of
    # the device-side record decoder tests
    shutil.copyfile(os.path.join(REFDIR, "sp1_dna.blow5"), os.path.join(HERE, "sp1_dna.blow5"))
    for name, src in (("nCoV-2019", "nCoV-2019.reference.fasta"), ("rnasequin", "rnasequin_sequences_2.4.fa")):
        with open(os.path.join(REFDIR, src), "rb") as fi, gzip.GzipFile(os.path.join(HERE, name + ".fa.gz"), "wb", mtime=0) as fo:
            shutil.copyfileobj(fi, fo)

    # 2. synthetic fixtures
    models = {k: synth.make_model(k) for k in (5, 6, 9)}
    ncov_names, ncov_seqs = H.read_fasta(os.path.join(HERE, "nCoV-2019.fa.gz"))
    seq_names, seq_seqs = H.read_fasta(os.path.join(HERE, "rnasequin.fa.gz"))

    sigs, _ = synth.simulate_reads(ncov_seqs, 6, models[6][0], 48, seed=11)
    save_reads(os.path.join(HERE, "synth_dna48.npz"), [f"synth_dna_{i:04d}" for i in range(48)], sigs,
               [synth.DNA_SCALING] * 48)

    rng = np.random.default_rng(5)
    multi = [synth.random_sequence(int(n), rng) for n in (5000, 1200, 23456, 777, 9000, 400)]
    # sprinkle non-ACGT and lower-case bases (ref.h:13-26, 45-65)
    m0 = bytearray(multi[0]); m0[100:104] = b"NNNN"; m0[2000:2010] = m0[2000:2010].lower(); multi[0] = bytes(m0)
    with gzip.GzipFile(os.path.join(HERE, "synth_multi.fa.gz"), "wb", mtime=0) as fo:
        for i, s in enumerate(multi):
            fo.write(f">contig{i} synthetic\n".encode())
            for o in range(0, len(s), 60):
                fo.write(s[o:o + 60] + b"\n")
    sigs, _ = synth.simulate_reads(multi, 6, models[6][0], 24, seed=12, bases_per_read=380)
    save_reads(os.path.join(HERE, "synth_dna_multi.npz"), [f"synth_multi_{i:04d}" for i in range(24)], sigs,
               [synth.DNA_SCALING] * 24)
    sigs, _ = synth.simulate_reads(multi, 9, models[9][0], 12, seed=15, bases_per_read=380)
    save_reads(os.path.join(HERE, "synth_dna_k9.npz"), [f"synth_k9_{i:04d}" for i in range(12)], sigs,
               [synth.DNA_SCALING] * 12)

    # reads of assorted lengths around the ignored / too-short thresholds (sigfish.c:450-461)
    sigs = []
    r2 = np.random.default_rng(13)
    for nb in (20, 45, 60, 70, 76, 80, 100, 150, 200, 260, 290, 310, 330, 400):
        s, _ = synth.simulate_reads(ncov_seqs, 6, models[6][0], 1, seed=int(r2.integers(1 << 30)), bases_per_read=nb,
                                    min_samples=800)
        sigs.append(s[0])
    save_reads(os.path.join(HERE, "synth_dna_short.npz"), [f"synth_short_{i:02d}" for i in range(len(sigs))], sigs,
               [synth.DNA_SCALING] * len(sigs))

    sigs, _ = synth.simulate_reads(seq_seqs, 5, models[5][0], 32, seed=14, rna=True, bases_per_read=420)
    save_reads(os.path.join(HERE, "synth_rna32.npz"), [f"synth_rna_{i:04d}" for i in range(32)], sigs,
               [synth.RNA_SCALING] * 32)

    sigs, _ = synth.simulate_rna_reads_with_tail(seq_seqs, 5, models[5][0], 24, seed=16)
    save_reads(os.path.join(HERE, "synth_rna_tail24.npz"), [f"synth_tail_{i:04d}" for i in range(24)], sigs,
               [synth.RNA_SCALING] * 24)

    tx_names, tx_seqs = H.case_fasta("gen_rna004_tx2000")
    sigs, _ = synth.simulate_reads(tx_seqs, 9, models[9][0], 24, seed=17, rna=True, bases_per_read=420)
    save_reads(os.path.join(HERE, "synth_rna004_24.npz"), [f"synth_rna004_{i:04d}" for i in range(24)], sigs,
               [synth.RNA_SCALING] * 24)
    sigs, _ = synth.simulate_rna_reads_with_tail(tx_seqs, 9, models[9][0], 16, seed=18, bases_per_read=500)
    save_reads(os.path.join(HERE, "synth_rna004_tail16.npz"), [f"synth_rna004_tail_{i:04d}" for i in range(16)], sigs,
               [synth.RNA_SCALING] * 16)

    # 3. golden PAFs from the reference binary
    for k, (mean, stdv) in models.items():
        synth.write_model_file(os.path.join(tmp, f"model_k{k}.txt"), k, mean, stdv)
    summary = {}
    for case, (reads, fasta, k, flags, q, p) in CASES.items():
        ids, sg, sc = H.load_reads_npz(os.path.join(HERE, reads + ".npz"))
        s5 = os.path.join(tmp, reads + ".slow5")
        rna = bool(flags & H.F_RNA)
        extra = CASE_EXTRA.get(case, {})
        s5 = os.path.join(tmp, reads + ("_" + extra["kit"] if extra else "") + ".slow5")
        synth.write_slow5_ascii(s5, ids, sg, rna=rna, kit=extra.get("kit"), scalings=sc)
        fa = os.path.join(tmp, fasta + ".fa")
        if fasta in H.GENERATED_FASTA:
            H.write_case_fasta(fasta, fa)
        else:
            with gzip.open(os.path.join(HERE, fasta + ".fa.gz"), "rb") as fi, open(fa, "wb") as fo:
                shutil.copyfileobj(fi, fo)
        paf = H.run_ref(fa, s5, os.path.join(tmp, f"model_k{k}.txt"), flags=flags, q=q, p=p)
        with open(os.path.join(HERE, "paf", case + ".paf"), "w") as f:
            f.write(paf)
        summary[case] = dict(reads=reads, fasta=fasta, k=k, flags=flags, q=q, p=p, rows=paf.count("\n"), **extra)
        if fasta in H.GENERATED_FASTA:
            import hashlib
            summary[case]["fasta_sha1"] = hashlib.sha1(b"\n".join(H.case_fasta(fasta)[1])).hexdigest()
        print(case, summary[case]["rows"], "rows")
    with open(os.path.join(HERE, "cases.json"), "w") as f:
        json.dump(summary, f, indent=1, sort_keys=True)

    # 3b. golden SAM (--sam) for a subset
    os.makedirs(os.path.join(HERE, "sam"), exist_ok=True)
    for case in SAM_CASES:
        reads, fasta, k, flags, q, p = CASES[case]
        extra = CASE_EXTRA.get(case, {})
        sam = H.run_ref(os.path.join(tmp, fasta + ".fa"), os.path.join(tmp, reads + ("_" + extra["kit"] if extra else "") + ".slow5"),
                        os.path.join(tmp, f"model_k{k}.txt"), flags=flags, q=q, p=p, extra=["--sam"])
        with open(os.path.join(HERE, "sam", case + ".sam"), "w") as f:
            f.write(sam)
        print("sam", case, sam.count("\n"), "lines")

    # 3c. `sigfish eval` reports (stdout) for truth/test pairs made of the golden PAFs plus one crafted pair
    #     with secondary records, several truth mappings per read and reads missing from the truth set
    import subprocess
    ev = os.path.join(HERE, "eval")
    os.makedirs(ev, exist_ok=True)
    rows = open(os.path.join(HERE, "paf", "dna_synth48.paf")).read().splitlines()
    crafted_truth, crafted_test = [], []
    for i, r in enumerate(rows[:30]):
        f = r.split("\t")
        crafted_truth.append("\t".join(f[:12] + ["tp:A:P"]))
        if i % 3 == 0:  # a secondary truth mapping somewhere else
            g = list(f)
            g[7], g[8] = str(int(f[7]) + 5000), str(int(f[8]) + 5000)
            crafted_truth.append("\t".join(g[:12] + ["tp:A:S"]))
        t = list(f)
        if i % 4 == 1:   # shifted by 5000: correct only against the secondary truth record
            t[7], t[8] = str(int(f[7]) + 5000 + 40), str(int(f[8]) + 5000 - 30)
        elif i % 4 == 2:  # wrong strand
            t[4] = "-" if f[4] == "+" else "+"
        elif i % 4 == 3:  # start off by 150, end off by 60: still correct (min of the two)
            t[7], t[8] = str(int(f[7]) + 150), str(int(f[8]) + 60)
        t[11] = str((i * 7) % 61)
        crafted_test.append("\t".join(t[:12] + ["tp:A:P"]))
    for r in rows[40:44]:  # reads the truth set does not have
        crafted_test.append(r)
    open(os.path.join(ev, "crafted_truth.paf"), "w").write("\n".join(crafted_truth) + "\n")
    open(os.path.join(ev, "crafted_test.paf"), "w").write("\n".join(crafted_test) + "\n")
    pafd = os.path.join(HERE, "paf")
    for name, (truth, test, opts) in EVAL_CASES.items():
        tp = os.path.join(ev, truth) if truth.startswith("crafted") else os.path.join(pafd, truth)
        sp = os.path.join(ev, test) if test.startswith("crafted") else os.path.join(pafd, test)
        r = subprocess.run([H.REF_BIN, "eval"] + opts + [tp, sp], capture_output=True, text=True, check=True)
        open(os.path.join(ev, name + ".txt"), "w").write(r.stdout)
        print("eval", name, r.stdout.count("\n"), "lines")

    # 4. event tables straight from the reference's getevents()
    for name, rna in (("sp1_dna", False), ("sequin_rna", True), ("synth_dna_short", False)):
        ids, sg, sc = H.load_reads_npz(os.path.join(HERE, name + ".npz"))
        evs = [ref_events(s, c, rna) for s, c in zip(sg, sc)]
        offs = np.zeros(len(evs) + 1, dtype=np.int64)
        offs[1:] = np.cumsum([len(e) for e in evs])
        cat = np.concatenate(evs)
        np.savez_compressed(os.path.join(HERE, f"events_{name}.npz"), offsets=offs, start=cat["start"],
                            length=cat["length"], mean=cat["mean"], stdv=cat["stdv"])
        print("events", name, [len(e) for e in evs])


if __name__ == "__main__":
    main()
