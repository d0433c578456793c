"""Pin the CPU oracle (oracle/*.c) against the unmodified reference.

 * golden PAFs under tests/golden/paf/ were produced by oracle/_ref/sigfish (the reference
   compiled from /root/reference, see tests/golden/make_golden.py);
 * golden event tables were produced by calling the reference's getevents() directly;
 * when oracle/_ref/libsigfish_ref.so is present the DTW kernels are also compared call by call.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

import helpers as H
from sigfish_b200 import synth

CASES = json.load(open(os.path.join(H.GOLDEN, "cases.json")))
MODELS = {}


def model(k):
    if k not in MODELS:
        MODELS[k] = synth.make_model(k)[0]
    return MODELS[k]


@pytest.mark.parametrize("case", sorted(CASES))
def test_oracle_paf_matches_reference_golden(case):
    c = CASES[case]
    names, seqs = H.case_fasta(c)
    ids, sigs, sc = H.case_reads(c)
    want = open(os.path.join(H.GOLDEN, "paf", case + ".paf")).read()
    assert want.count("\n") == c["rows"]
    n = c.get("oracle_reads")
    if n:  # a heavy case: the first n reads only (every read of such a case maps, so they are the first n lines)
        assert c["rows"] == len(ids)
        ids, sigs, sc = ids[:n], sigs[:n], sc[:n]
        want = "".join(want.splitlines(keepends=True)[:n])
    got = H.oracle_paf(names, seqs, model(c["k"]), c["k"], ids, sigs, sc, H.case_oracle_flags(c), c["q"], c["p"])
    assert got == want


@pytest.mark.parametrize("name,rna", [("sp1_dna", False), ("sequin_rna", True), ("synth_dna_short", False)])
def test_oracle_events_match_reference_getevents(name, rna):
    ids, sigs, sc = H.load_reads_npz(os.path.join(H.GOLDEN, name + ".npz"))
    z = np.load(os.path.join(H.GOLDEN, f"events_{name}.npz"))
    offs = z["offsets"]
    for i, (s, c) in enumerate(zip(sigs, sc)):
        ev = H.orc_events(s, c["digitisation"], c["offset"], c["range"], rna)
        a, b = offs[i], offs[i + 1]
        assert len(ev) == b - a
        assert np.array_equal(ev["start"], z["start"][a:b])
        # bit-exact fp32
        for f in ("length", "mean", "stdv"):
            assert np.array_equal(ev[f].view(np.uint32), z[f][a:b].view(np.uint32)), (name, i, f)


def _ref_lib():
    if not os.path.exists(H.REF_SO):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    L = C.CDLL(H.REF_SO)
    f32 = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
    L.subsequence.argtypes = [f32, f32, C.c_int, C.c_int, f32]
    L.std_dtw.argtypes = [f32, f32, C.c_int, C.c_int, f32, C.c_int]
    L.std_dtw.restype = C.c_float

    class Path(C.Structure):
        _fields_ = [("k", C.c_int), ("px", C.POINTER(C.c_int)), ("py", C.POINTER(C.c_int))]
    L.subsequence_path.argtypes = [f32, C.c_int, C.c_int, C.c_int, C.POINTER(Path)]
    L.subsequence_path.restype = C.c_int
    L.Path = Path
    return L


@pytest.mark.refbin
@pytest.mark.parametrize("quant", [0, 4, 1])
def test_oracle_dtw_matches_reference_functions(quant):
    """random and tie-heavy (quantised) inputs: cost matrices, path start and full path"""
    R = _ref_lib()
    O = H.oracle()
    rng = np.random.default_rng(100 + quant)
    libc = C.CDLL(None)
    for trial in range(60):
        n = int(rng.integers(1, 40))
        m = int(rng.integers(1, 120))
        x = rng.normal(size=n).astype(np.float32)
        y = rng.normal(size=m).astype(np.float32)
        if quant:
            x = (np.round(x * quant) / quant).astype(np.float32)
            y = (np.round(y * quant) / quant).astype(np.float32)
        for std in (False, True):
            ca = np.zeros(n * m, dtype=np.float32)
            cb = np.zeros(n * m, dtype=np.float32)
            if std:
                ra = R.std_dtw(x, y, n, m, ca, 0)
                rb = O.orc_std_dtw(x, y, n, m, cb)
                assert np.float32(ra).view(np.uint32) == np.float32(rb).view(np.uint32)
            else:
                R.subsequence(x, y, n, m, ca)
                O.orc_subsequence(x, y, n, m, cb)
            assert np.array_equal(ca.view(np.uint32), cb.view(np.uint32))
            for end in {m - 1, int(rng.integers(0, m)), 0}:
                p = R.Path()
                assert R.subsequence_path(ca, n, m, end, C.byref(p)) == 1
                ref_px = np.ctypeslib.as_array(p.px, shape=(p.k,)).copy()
                ref_py = np.ctypeslib.as_array(p.py, shape=(p.k,)).copy()
                libc.free(p.px)
                libc.free(p.py)
                assert O.orc_path_start(cb, n, m, end) == ref_py[0]
                px = np.zeros(n + m, dtype=np.int32)
                py = np.zeros(n + m, dtype=np.int32)
                k = O.orc_path_full(cb, n, m, end, px, py)
                assert k == p.k
                assert np.array_equal(px[:k], ref_px) and np.array_equal(py[:k], ref_py)


@pytest.mark.refbin
def test_oracle_paf_matches_reference_binary_fresh_inputs(tmp_path):
    """new seeded inputs, run through the reference binary right now (container only)"""
    if not H.have_ref_bin():
        pytest.skip("oracle/_ref not built")
    k = 6
    mean, stdv = synth.make_model(k, seed=21)
    rng = np.random.default_rng(77)
    seqs = [synth.random_sequence(int(n), rng) for n in (3000, 800, 6000)]
    names = [f"c{i}" for i in range(len(seqs))]
    sigs, _ = synth.simulate_reads(seqs, k, mean, 10, seed=78, bases_per_read=400)
    ids = [f"r{i}" for i in range(len(sigs))]
    fa = str(tmp_path / "ref.fa")
    s5 = str(tmp_path / "reads.slow5")
    mf = str(tmp_path / "model.txt")
    synth.write_fasta(fa, names, seqs)
    synth.write_slow5_ascii(s5, ids, sigs)
    synth.write_model_file(mf, k, mean, stdv)
    for flags, q, p in ((0, 250, 50), (H.F_END, 250, 50), (0, 64, 10)):
        want = H.run_ref(fa, s5, mf, flags=flags, q=q, p=p)
        got = H.oracle_paf(names, seqs, mean, k, ids, sigs, [synth.DNA_SCALING] * len(sigs), flags, q, p)
        assert got == want


def test_top_list_later_equal_score_wins():
    """equal scores -- the later candidate ranks better (sigfish.c:577-583).
    Two identical contigs give identical candidate scores; the reference reports the later one."""
    k = 6
    mean = model(k)
    rng = np.random.default_rng(3)
    s = synth.random_sequence(1500, rng)
    sigs, _ = synth.simulate_reads([s], k, mean, 3, seed=4, bases_per_read=400, both_strands=False)
    ref = H.OracleRef([s, s], mean, k, 0, 250)
    for sig in sigs:
        hit = H.orc_map(ref, sig, 8192.0, 10.0, 1402.882, 0, 250, 50)
        assert hit.mapped and hit.rid == 1 and hit.score == hit.score2 and hit.mapq == 0


SAM_CASES = sorted(f[:-4] for f in os.listdir(os.path.join(H.GOLDEN, "sam")))


@pytest.mark.parametrize("case", SAM_CASES)
def test_oracle_sam_matches_reference_golden(case):
    """--sam: the winner's full warping path turned into the ss:Z string (sigfish.c:530-571, 663-794)"""
    c = CASES[case]
    names, seqs = H.case_fasta(c)
    ids, sigs, sc = H.case_reads(c)
    want = open(os.path.join(H.GOLDEN, "sam", case + ".sam")).read()
    want = [l for l in want.splitlines(keepends=True) if not l.startswith("@PG")]
    n = c.get("oracle_reads") or (8 if case.startswith("rna004_") else 0)
    if n:  # heavy cases: header + the first n records (every read of such a case has a record)
        ids, sigs, sc = ids[:n], sigs[:n], sc[:n]
        want = want[:len(seqs) + n]
    got = H.oracle_sam(names, seqs, model(c["k"]), c["k"], ids, sigs, sc, H.case_oracle_flags(c), c["q"], c["p"])
    assert got == "".join(want)


def _fuzz_seeds():
    """16 seeds in the suite; SF_FUZZ_SEEDS=first:last runs another range (long runs are recorded in DESIGN.md 3)"""
    r = os.environ.get("SF_FUZZ_SEEDS", "")
    if ":" in r:
        a, b = r.split(":")
        return range(int(a), int(b))
    return range(16)


@pytest.mark.refbin
@pytest.mark.parametrize("seed", _fuzz_seeds())
def test_oracle_fuzz_against_reference_binary(tmp_path, seed):
    """seeded fuzz of the whole path, oracle vs the unmodified reference binary run right now (build container
    only): chemistry, every flag combination the CLI accepts, q, p (incl. -1), contig counts / lengths, PAF and SAM"""
    if not H.have_ref_bin():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(7000 + seed)
    rna = bool(rng.integers(0, 2))
    k = (9 if seed % 3 == 2 else 5) if rna else int(rng.choice([6, 9]))  # RNA: r9 5-mers and RNA004 9-mers
    rna004 = rna and k == 9
    flags = 0
    p = int(rng.choice([0, 10, 50, 50, 120]))
    if rna:
        flags = H.F_RNA | int(rng.choice([0, H.F_DTW, H.F_INV, H.F_REF, H.F_DTW | H.F_REF, H.F_INV | H.F_REF]))
        if rng.integers(0, 3) == 0 and not (flags & H.F_INV):
            p = -1
    if p >= 0 and rng.integers(0, 3) == 0:
        flags |= H.F_END
    q = int(rng.choice([25, 60, 97, 130, 250, 250, 333]))
    mean, stdv = synth.make_model(k, seed=31 + seed)
    seqs = [synth.random_sequence(int(n), rng) for n in rng.integers(700, 4000, size=int(rng.integers(1, 5)))]
    names = [f"c{i}" for i in range(len(seqs))]
    sigs, scs = [], []
    for r in range(6):
        if rna and p < 0 and r % 2 == 0:
            s, _ = synth.simulate_rna_reads_with_tail(seqs, k, mean, 1, seed=int(rng.integers(1 << 30)), bases_per_read=500)
        else:
            s, _ = synth.simulate_reads(seqs, k, mean, 1, seed=int(rng.integers(1 << 30)), rna=rna,
                                        bases_per_read=int(rng.choice([150, 300, 450, 700])), min_samples=700)
        sigs.append(s[0])
        scs.append(synth.RNA_SCALING if rna else synth.DNA_SCALING)
    ids = [f"r{i}" for i in range(len(sigs))]
    fa, s5, mf = str(tmp_path / "ref.fa"), str(tmp_path / "reads.blow5"), str(tmp_path / "model.txt")
    synth.write_fasta(fa, names, seqs)
    synth.write_blow5(s5, ids, sigs, rna=rna, kit="sqk-rna004" if rna004 else ("sqk-lsk114" if k == 9 else None), scalings=scs)
    synth.write_model_file(mf, k, mean, stdv)
    oflags = flags | (0x400 if rna004 else 0)  # the reference picks the rna004 jnn parameters from the kit
    want = H.run_ref(fa, s5, mf, flags=flags, q=q, p=p)
    got = H.oracle_paf(names, seqs, mean, k, ids, sigs, scs, oflags, q, p)
    assert got == want, (flags, q, p)
    if not flags & H.F_DTW:  # the reference aborts on --dtw-std --sam (sigfish.c:669)
        want = H.run_ref(fa, s5, mf, flags=flags, q=q, p=p, extra=["--sam"])
        want = "".join(l for l in want.splitlines(keepends=True) if not l.startswith("@PG"))
        assert H.oracle_sam(names, seqs, mean, k, ids, sigs, scs, oflags, q, p) == want, (flags, q, p)
