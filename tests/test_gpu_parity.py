"""GPU parity tests: the CUDA path, called through the C-ABI (libsfgpu.so), against the CPU oracle
and the golden vectors produced by the unmodified reference.

Bar: bit-exact.  Integer fields (coordinates, strand, contig, window) must be equal; fp32 scores
are compared by bit pattern (the north star allows 1e-6 relative, we require 0 ulp).
"""
import json
import os

import numpy as np
import pytest

import helpers as H
from sigfish_b200 import capi, synth

pytestmark = pytest.mark.gpu

CASES = json.load(open(os.path.join(H.GOLDEN, "cases.json")))
_MODELS = {}


def model(k):
    if k not in _MODELS:
        _MODELS[k] = synth.make_model(k)[0]
    return _MODELS[k]


def bits(x):
    return np.asarray(x, dtype=np.float32).view(np.uint32)


def assert_hit_equal(g, o, tag, flags, q, p):
    """g: one row of capi.RESULT_DTYPE, o: helpers.OrcHit"""
    if not o.mapped:
        assert g["qlen"] == 0, tag
        return
    assert g["qlen"] == o.qend - o.qstart, tag
    assert (g["qstart"], g["qend"]) == (o.qstart, o.qend), tag
    assert (g["status"] & 3) == (o.status & 3), tag
    assert bool(g["status"] & 16) == bool(o.status & 4), tag  # automatic query start failed
    assert (int(g["start_raw"]), int(g["end_raw"])) == (o.start_raw, o.end_raw), tag
    if flags & H.F_END:
        assert g["n_events"] == o.n_events, tag
    elif p < 0:
        assert g["qend"] < g["n_events"] <= o.n_events or g["n_events"] == o.n_events, tag
    else:
        assert g["n_events"] == min(o.n_events, p + q + 1) or g["n_events"] == o.n_events, tag
    assert g["rid"] == o.rid, tag
    assert "+-"[g["strand"]] == o.strand.decode(), tag
    assert bits(g["score"]) == bits(o.score), (tag, float(g["score"]), o.score)
    assert bits(g["score2"]) == bits(o.score2), (tag, float(g["score2"]), o.score2)
    assert (g["pos_st"], g["pos_end"]) == (o.raw_pos_st, o.raw_pos_end), tag


# ------------------------------------------------------------------ kernel #2: reference synthesis

REF_COMBOS = sorted({(c["fasta"], c["k"], c["flags"] & ~H.F_DTW, c["q"]) for c in CASES.values()})


@pytest.mark.parametrize("fasta,k,flags,q", REF_COMBOS)
def test_ref_events_bit_exact(fasta, k, flags, q):
    names, seqs = H.case_fasta(fasta)
    ctx = capi.Context(model(k), k, flags=flags, query_size=q)
    ctx.set_ref(seqs)
    ref = H.OracleRef(seqs, model(k), k, flags, q)
    for i in range(len(seqs)):
        assert ctx.ref_lengths[i] == ref.length(i)
        assert ctx.ref_st_offset[i] == ref.offset(i)
        assert ctx.ref_seq_lengths[i] == len(seqs[i])
        assert np.array_equal(bits(ctx.ref_events(i, 0)), bits(ref.fwd(i))), (fasta, i, "+")
        if not flags & H.F_RNA:
            assert np.array_equal(bits(ctx.ref_events(i, 1)), bits(ref.rev(i))), (fasta, i, "-")
    ref.close()
    ctx.close()


def test_ref_non_acgt_and_lowercase():
    k = 6
    rng = np.random.default_rng(5)
    s = bytearray(synth.random_sequence(3000, rng))
    for i in rng.integers(0, 3000, size=60):
        s[i] = ord("N")
    for i in rng.integers(0, 3000, size=200):
        s[i] = ord(chr(s[i]).lower())
    seqs = [bytes(s), synth.random_sequence(k, rng), synth.random_sequence(k + 1, rng)]
    ctx = capi.Context(model(k), k)
    ctx.set_ref(seqs)
    ref = H.OracleRef(seqs, model(k), k, 0, 250)
    for i in (0, 2):  # a 1-column contig z-scores to NaN on both paths; compare the two others
        assert np.array_equal(bits(ctx.ref_events(i, 0)), bits(ref.fwd(i)))
        assert np.array_equal(bits(ctx.ref_events(i, 1)), bits(ref.rev(i)))
    assert ctx.ref_lengths.tolist() == [2995, 1, 2]


# ------------------------------------------------------------------ kernel #1: event detection

@pytest.mark.parametrize("name,rna", [("sp1_dna", False), ("sequin_rna", True), ("synth_dna_short", False)])
def test_event_tables_match_reference_golden(name, rna):
    ids, sigs, sc = H.load_reads_npz(os.path.join(H.GOLDEN, name + ".npz"))
    z = np.load(os.path.join(H.GOLDEN, f"events_{name}.npz"))
    offs = z["offsets"]
    k = 5 if rna else 6
    ctx = capi.Context(model(k), k, flags=H.F_RNA if rna else 0)
    for i, (s, c) in enumerate(zip(sigs, sc)):
        start, length, mean = ctx.event_table(s, c)
        a, b = offs[i], offs[i + 1]
        assert len(start) == b - a, (name, i)
        assert np.array_equal(start, z["start"][a:b]), (name, i)
        assert np.array_equal(bits(length), bits(z["length"][a:b])), (name, i)
        assert np.array_equal(bits(mean), bits(z["mean"][a:b])), (name, i)
    ctx.close()


def test_event_tables_random_scalings():
    """reads whose pA values are not on a coarse grid: prefix sums must still equal the sequential ones"""
    rng = np.random.default_rng(11)
    k = 6
    ctx = capi.Context(model(k), k)
    for trial in range(6):
        n = int(rng.integers(400, 9000))
        lv = rng.uniform(60, 130, size=n // 8 + 2).astype(np.float32)
        sc = dict(digitisation=float(rng.choice([8192.0, 2048.0, 1000.0])), range=float(rng.uniform(400, 1500)),
                  offset=float(rng.uniform(-300, 50)), sampling_rate=4000.0)
        sig = synth.simulate_read(lv, rng, sc)[:n]
        ev = H.orc_events(sig, sc["digitisation"], sc["offset"], sc["range"], False)
        start, length, mean = ctx.event_table(sig, sc)
        assert np.array_equal(start, ev["start"]), trial
        assert np.array_equal(bits(length), bits(ev["length"])), trial
        assert np.array_equal(bits(mean), bits(ev["mean"])), trial
    ctx.close()


# ------------------------------------------------------------------ whole path on the golden cases

@pytest.mark.parametrize("case", sorted(CASES))
def test_mapping_matches_oracle_on_golden_cases(case):
    c = CASES[case]
    names, seqs = H.case_fasta(c)
    ids, sigs, sc = H.case_reads(c)
    k, flags, q, p = c["k"], c["flags"], c["q"], c["p"]
    ref = H.OracleRef(seqs, model(k), k, flags, q)
    oflags = H.case_oracle_flags(c)
    n_orc = c.get("oracle_reads") or len(sigs)  # heavy cases: the oracle checks the first reads only
    want = [H.orc_map(ref, s, cc["digitisation"], cc["offset"], cc["range"], oflags, q, p)
            for s, cc in list(zip(sigs, sc))[:n_orc]]
    if n_orc == len(sigs):
        assert sum(int(o.mapped) for o in want) == c["rows"]
    ctx = capi.Context(model(k), k, flags=flags, query_size=q, prefix_size=p, pore=c.get("pore", 0))
    ctx.set_ref(seqs)
    got = ctx.map_batch(sigs, sc)
    for i, o in enumerate(want):
        assert_hit_equal(got[i], o, (case, i), flags, q, p)
    ctx.close()
    ref.close()


def test_batch_slots_and_resubmit_are_deterministic():
    c = CASES["dna_synth48"]
    names, seqs = H.case_fasta(c)
    ids, sigs, sc = H.case_reads(c)
    ctx = capi.Context(model(c["k"]), c["k"], n_slots=2)
    ctx.set_ref(seqs)
    half = len(sigs) // 2
    ctx.submit(0, *ctx.pack(sigs[:half], sc[:half]))
    ctx.submit(1, *ctx.pack(sigs[half:], sc[half:]))
    a = np.concatenate([ctx.collect(0), ctx.collect(1)])
    b = ctx.map_batch(sigs, sc, slot=0)
    assert a.tobytes() == b.tobytes()
    ctx.resubmit(0)
    assert ctx.collect(0).tobytes() == b.tobytes()
    ctx.submit_reads(1, sigs, sc)  # one pointer per read instead of one flat buffer
    assert ctx.collect(1).tobytes() == b.tobytes()
    t = ctx.timing(0)
    assert 1 <= t.dtw_launches <= 4 and t.cells > 0 and t.dtw_ms > 0  # pair kernel + warp-per-read kernel (+ their redo passes)
    ctx.close()


def test_empty_batch_and_empty_read():
    k = 6
    rng = np.random.default_rng(2)
    seqs = [synth.random_sequence(2000, rng)]
    ctx = capi.Context(model(k), k)
    ctx.set_ref(seqs)
    assert len(ctx.map_batch([], [])) == 0
    sigs, _ = synth.simulate_reads(seqs, k, model(k), 2, seed=3, bases_per_read=400)
    out = ctx.map_batch([sigs[0], np.zeros(0, np.int16), sigs[1]], [synth.DNA_SCALING] * 3)
    assert out["qlen"].tolist()[1] == 0 and out["qlen"][0] > 0 and out["qlen"][2] > 0
    ref = H.OracleRef(seqs, model(k), k, 0, 250)
    for j, i in ((0, 0), (2, 1)):
        o = H.orc_map(ref, sigs[i], 8192.0, 10.0, 1402.882, 0, 250, 50)
        assert_hit_equal(out[j], o, j, 0, 250, 50)
    ctx.close()


# ------------------------------------------------------------------ kernel #3 alone: arbitrary, tie-heavy inputs

def _rand_arrays(rng, lens, quant):
    out = []
    for n in lens:
        y = rng.normal(size=n).astype(np.float32)
        if quant:
            y = (np.round(y * quant) / quant).astype(np.float32)
        out.append(y)
    return out


@pytest.mark.parametrize("q", [20, 33, 64, 100, 250, 300, 500])
@pytest.mark.parametrize("quant", [0, 2, 1])
def test_dtw_kernel_matches_oracle_random_and_ties(q, quant):
    """random and quantised (tie-heavy) queries / references, both strands, ragged query lengths;
    checkpointing forced on (ck_min_cols) with a tiny restart window so that the start-coordinate pass
    takes the exit-code path"""
    rng = np.random.default_rng(1000 * q + quant)
    lens = [int(x) for x in rng.integers(1, 900, size=5)] + [3000, 1, q, q + 1]
    fwd = _rand_arrays(rng, lens, quant)
    rev = _rand_arrays(rng, lens, quant)
    qlens = [q, q, max(1, q - 1), max(1, q // 2), 25 if q > 25 else q, 1, q, q]
    queries = _rand_arrays(rng, qlens, quant)
    oref = H.OracleEventRef(fwd, rev)
    for ck, win in ((0, 0), (256, 16), (128, 1)):
        ctx = capi.Context(model(5), 5, query_size=q, ck_min_cols=ck, min_window=win)
        ctx.set_ref_events(fwd, rev)
        got = ctx.align_queries(queries)
        for i, x in enumerate(queries):
            o = oref.align(x, 0)
            g = got[i]
            tag = (q, quant, ck, i)
            assert g["rid"] == o.rid, tag
            assert "+-"[g["strand"]] == o.strand.decode(), tag
            assert bits(g["score"]) == bits(o.score), tag
            assert bits(g["score2"]) == bits(o.score2), tag
            assert (g["pos_st"], g["pos_end"]) == (o.raw_pos_st, o.raw_pos_end), tag
        ctx.close()
    oref.close()


@pytest.mark.parametrize("q", [100, 180, 250, 256, 300, 500])
@pytest.mark.parametrize("quant", [0, 2])
@pytest.mark.parametrize("warm", [0, 1])
def test_split_segments_match_oracle(q, quant, warm):
    """long segments cut into column pieces: every piece warms up behind a +INF column, its front
    at the piece boundary is compared with its predecessor's, differing pieces are redone.  warm=0: the default
    warm-up of 2q columns; warm=1: one block (63 columns < q), so that most fronts differ and the redo pass does
    the work.  Random and tie-heavy inputs, ragged query lengths, queries cut out of the reference across piece
    boundaries (score 0 hits whose chunk is shared by two pieces)."""
    rng = np.random.default_rng(4000 * q + 10 * quant + warm)
    lens = [20000, 700, 9000, 64 * 40 + 1, 5]
    fwd = _rand_arrays(rng, lens, quant)
    rev = _rand_arrays(rng, lens, quant)
    qlens = [q, q, q, max(1, q - 1), max(1, q // 2), 25 if q > 25 else q, q, q]
    queries = _rand_arrays(rng, qlens, quant)
    for st in range(0, 20000 - q, 331):  # planted exact matches all along the long segment
        src = fwd[0] if (st // 331) % 2 == 0 else rev[0]
        queries.append(src[st:st + q].copy())
    oref = H.OracleEventRef(fwd, rev)
    want = [oref.align(x, 0) for x in queries]
    for periods in (1, 3):
        ctx = capi.Context(model(5), 5, query_size=q, ck_min_cols=128, min_window=16, warm_blocks=warm,
                           piece_periods=periods)
        ctx.set_ref_events(fwd, rev)
        got = ctx.align_queries(queries)
        t = ctx.timing(0)
        assert t.piece_blocks > 0 and t.tasks_per_read > 2 * len(lens), (t.piece_blocks, t.tasks_per_read)
        if warm == 1:
            assert t.redone_pieces > 0
        for i, o in enumerate(want):
            g = got[i]
            tag = (q, quant, warm, periods, i)
            assert g["rid"] == o.rid, tag
            assert "+-"[g["strand"]] == o.strand.decode(), tag
            assert bits(g["score"]) == bits(o.score), tag
            assert bits(g["score2"]) == bits(o.score2), tag
            assert (g["pos_st"], g["pos_end"]) == (o.raw_pos_st, o.raw_pos_end), tag
        # the same batch with splitting switched off gives the same bytes
        ctx0 = capi.Context(model(5), 5, query_size=q, ck_min_cols=128, min_window=16, piece_periods=-1)
        ctx0.set_ref_events(fwd, rev)
        got0 = ctx0.align_queries(queries)
        assert ctx0.timing(0).piece_blocks == 0
        assert got0.tobytes() == got.tobytes()
        ctx0.close()
        ctx.close()
    oref.close()


def test_split_reads_through_whole_path_and_sam():
    """reads (events + DTW + start coordinate + --sam paths) on a reference whose contigs are split into pieces,
    forced redo included, against the oracle"""
    k = 6
    rng = np.random.default_rng(31)
    seqs = [synth.random_sequence(30000, rng), synth.random_sequence(2500, rng)]
    sigs, _ = synth.simulate_reads(seqs, k, model(k), 40, seed=32, bases_per_read=430)
    sigs += synth.simulate_reads(seqs, k, model(k), 6, seed=33, bases_per_read=150, min_samples=700)[0]  # ragged queries
    sc = [synth.DNA_SCALING] * len(sigs)
    ref = H.OracleRef(seqs, model(k), k, 0, 250)
    want = [H.orc_map(ref, s, c["digitisation"], c["offset"], c["range"], 0, 250, 50) for s, c in zip(sigs, sc)]
    base = None
    for warm, periods in ((0, 1), (1, 2), (0, 0)):
        ctx = capi.Context(model(k), k, flags=capi.SFGPU_SAM, ck_min_cols=256, warm_blocks=warm, piece_periods=periods)
        ctx.set_ref(seqs)
        got = ctx.map_batch(sigs, sc)
        t = ctx.timing(0)
        assert t.piece_blocks > 0
        if warm == 1:
            assert t.redone_pieces > 0
        for i, o in enumerate(want):
            assert_hit_equal(got[i], o, ("split", warm, periods, i), 0, 250, 50)
        paths, _, _ = ctx.collect_paths(0, got)
        blob = got.tobytes() + b"".join(b"-" if p is None else p[0].tobytes() + p[1].tobytes() for p in paths)
        base = base or blob
        assert blob == base
        ctx.close()
    ref.close()


@pytest.mark.parametrize("q", [30, 250, 375])
@pytest.mark.parametrize("quant", [0, 2])
def test_std_dtw_kernel_matches_oracle(q, quant):
    rng = np.random.default_rng(77 * q + quant)
    lens = [int(x) for x in rng.integers(1, 600, size=40)] + [1, 2, 2000]
    fwd = _rand_arrays(rng, lens, quant)
    qlens = [q, q - 1, q // 2, 25, 1, q]
    queries = _rand_arrays(rng, qlens, quant)
    oref = H.OracleEventRef(fwd, None)
    for ck, win in ((0, 0), (128, 1)):
        ctx = capi.Context(model(5), 5, flags=H.F_RNA | H.F_DTW, query_size=q, ck_min_cols=ck, min_window=win)
        ctx.set_ref_events(fwd, None)
        got = ctx.align_queries(queries)
        for i, x in enumerate(queries):
            # the ABI takes the query as dtw_single builds it; with RNA and no --invert the oracle
            # reverses its input, so hand it the reversed array
            o = oref.align(x[::-1].copy(), H.F_RNA | H.F_DTW)
            g = got[i]
            tag = (q, quant, ck, i)
            assert g["rid"] == o.rid and g["strand"] == 0, tag
            assert bits(g["score"]) == bits(o.score), tag
            assert bits(g["score2"]) == bits(o.score2), tag
            assert (g["pos_st"], g["pos_end"]) == (o.raw_pos_st, o.raw_pos_end), tag
        ctx.close()
    oref.close()


def test_equal_scores_later_candidate_wins():
    """identical contigs give identical candidate scores: the later one must be reported"""
    rng = np.random.default_rng(9)
    y = rng.normal(size=700).astype(np.float32)
    x = rng.normal(size=250).astype(np.float32)
    oref = H.OracleEventRef([y, y, y], [y, y, y])
    ctx = capi.Context(model(5), 5, query_size=250)
    ctx.set_ref_events([y, y, y], [y, y, y])
    g = ctx.align_queries([x])[0]
    o = oref.align(x, 0)
    assert (g["rid"], "+-"[g["strand"]]) == (2, "-") == (o.rid, o.strand.decode())
    assert bits(g["score"]) == bits(g["score2"]) == bits(o.score)
    assert (g["pos_st"], g["pos_end"]) == (o.raw_pos_st, o.raw_pos_end)
    ctx.close()


def test_long_reference_round_trip_property():
    """size-independent property at a length the CPU matrix cannot reach quickly: a query cut out of
    a 2 M-column reference must come back at its own coordinates with score 0, from either strand"""
    rng = np.random.default_rng(4)
    n = 2_000_000
    y = rng.normal(size=n).astype(np.float32)
    r = rng.normal(size=n).astype(np.float32)
    ctx = capi.Context(model(5), 5, query_size=250)
    ctx.set_ref_events([y], [r])
    starts = [0, 12345, 999_999, n - 250]
    queries = [y[s:s + 250] for s in starts] + [r[s:s + 250] for s in starts]
    got = ctx.align_queries(queries)
    for i, s in enumerate(starts * 2):
        g = got[i]
        assert g["score"] == 0.0 and g["strand"] == i // 4
        assert (g["pos_st"], g["pos_end"]) == (s, s + 249)
    ctx.close()


def test_100mb_reference_builds_fast_and_matches_oracle():
    """a 100 Mb contig (both strands: 2e8 reference columns): the z-score needs two serial fp32 sums over 1e8 values
    (genref.c:23-47; the running sum stagnates long before the end, which parity has to reproduce); the device
    does them at one FADD latency per element.  Events bit-exact against the oracle, queries cut from the events
    come back at their own coordinates through the split-segment path"""
    import time
    k = 6
    n = 100_000_000
    seq = synth.random_sequence(n, np.random.default_rng(123))
    ctx = capi.Context(model(k), k, query_size=250)
    t0 = time.perf_counter()
    ctx.set_ref([seq])
    dt = time.perf_counter() - t0
    ref = H.OracleRef([seq], model(k), k, 0, 250)
    fwd, rev = ref.fwd(0), ref.rev(0)
    ref.close()
    assert np.array_equal(bits(ctx.ref_events(0, 0)), bits(fwd))
    assert np.array_equal(bits(ctx.ref_events(0, 1)), bits(rev))
    assert dt < 3.0, dt  # includes the 100 MB host-to-device copy of the bases; the statistics kernel itself is ~0.5 s
    starts = [0, 31_000_123, 99_999_745]
    queries = [fwd[s:s + 250].copy() for s in starts] + [rev[s:s + 250].copy() for s in starts]
    got = ctx.align_queries(queries)
    t = ctx.timing(0)
    assert t.piece_blocks > 0 and t.redone_pieces == 0
    for i, s0 in enumerate(starts * 2):
        g = got[i]
        assert g["score"] == 0.0 and g["strand"] == i // 3, i
        assert (g["pos_st"], g["pos_end"]) == (s0, s0 + 249), i
    ctx.close()


def test_auto_query_start_rna004_parameter_set():
    """-p -1 with pore_flag == rna004 switches the adaptor finder to jnn.h:91-97 (std_scale 0.7, shortest dip 500)"""
    c = CASES["rna_tail24_auto"]
    names, seqs = H.case_fasta(c)
    ids, sigs, sc = H.case_reads(c)
    # shorten some adaptors so that the two parameter sets disagree
    sigs = [s[1800:] if i % 3 == 0 else s for i, s in enumerate(sigs)]
    k, q = c["k"], c["q"]
    ctx = capi.Context(model(k), k, flags=H.F_RNA, query_size=q, prefix_size=-1, pore=2)
    ctx.set_ref(seqs)
    got = ctx.map_batch(sigs, sc)
    ref = H.OracleRef(seqs, model(k), k, H.F_RNA, q)
    differs = 0
    for i, (s, cc) in enumerate(zip(sigs, sc)):
        o = H.orc_map(ref, s, cc["digitisation"], cc["offset"], cc["range"], H.F_RNA | 0x400, q, -1)
        assert_hit_equal(got[i], o, ("rna004", i), H.F_RNA, q, -1)
        o9 = H.orc_map(ref, s, cc["digitisation"], cc["offset"], cc["range"], H.F_RNA, q, -1)
        differs += int((o9.qstart, o9.status) != (o.qstart, o.status))
    ref.close()
    ctx.close()


@pytest.mark.parametrize("q", [33, 90, 200, 250, 300, 600])
@pytest.mark.parametrize("quant", [0, 2])
def test_warping_paths_match_oracle_backtrack(q, quant):
    """--sam support: sfgpu_collect_paths() must return exactly the path subsequence_path() finds in the full
    cost matrix (cdtw.c:98-167, 192-227), ties included; checked through real reads so that the event kernel
    also fills the window boundaries"""
    c = CASES["dna_multi_contig"]
    names, seqs = H.case_fasta(c)
    ids, sigs, sc = H.case_reads(c)
    k = c["k"]
    lm = model(k)
    if quant:  # a coarse model makes reference events repeat: many exact ties in the matrix
        lm = (np.round(lm / 8.0) * 8.0).astype(np.float32)
    ctx = capi.Context(lm, k, flags=capi.SFGPU_SAM, query_size=q, prefix_size=20)
    ctx.set_ref(seqs)
    got = ctx.map_batch(sigs, sc)
    paths, ev_start, ev_len = ctx.collect_paths(0, got)
    ref = H.OracleRef(seqs, lm, k, 0, q)
    O = H.oracle()
    for i, (s, cc) in enumerate(zip(sigs, sc)):
        o = H.orc_map(ref, s, cc["digitisation"], cc["offset"], cc["range"], 0, q, 20)
        assert_hit_equal(got[i], o, ("path", q, quant, i), 0, q, 20)
        ev = H.orc_events(s, cc["digitisation"], cc["offset"], cc["range"], False)
        lo, hi = o.qstart, o.qend
        assert np.array_equal(ev_start[i][:hi - lo], ev["start"][lo:hi])
        assert np.array_equal(bits(ev_len[i][:hi - lo]), bits(ev["length"][lo:hi]))
        x = ctx.query(0, i)
        y = ref.fwd(o.rid) if o.strand == b"+" else ref.rev(o.rid)
        n, m = len(x), len(y)
        cost = np.zeros(n * m, dtype=np.float32)
        O.orc_subsequence(x, y, n, m, cost)
        px = np.zeros(n + m, dtype=np.int32)
        py = np.zeros(n + m, dtype=np.int32)
        kk = O.orc_path_full(cost, n, m, o.raw_pos_end, px, py)
        assert paths[i] is not None
        assert np.array_equal(paths[i][0], px[:kk]) and np.array_equal(paths[i][1], py[:kk]), (q, quant, i)
    ref.close()
    ctx.close()


def test_c4_shaped_sample_truth_and_oracle_parity():
    """BASELINE.json configs[3] shape (R10 k=9, one 1 Mb contig, both strands, production checkpoint spacing):
    synthetic reads must come back at their true locus (the reference's own 85 % gate), the batch result must
    not depend on batch composition, and a few reads are checked field by field against the CPU oracle
    (3.4 s of CPU each: 2 x 250 x 999 992 cells)."""
    k = 9
    lm = model(k)
    rng = np.random.default_rng(1)
    seq = synth.random_sequence(1_000_000, rng)
    sigs, truth = synth.simulate_reads([seq], k, lm, 192, seed=77, bases_per_read=450)
    sc = [synth.DNA_SCALING] * len(sigs)
    ctx = capi.Context(lm, k)
    ctx.set_ref([seq])
    got = ctx.map_batch(sigs, sc)
    ok = 0
    for g, (ci, strand, st) in zip(got, truth):
        assert g["qlen"] == 250 and g["rid"] == 0
        # truth and the raw result are both in the coordinates of the strand's own event array;
        # the query starts ~50 events (~bases) into the read
        near = abs(int(g["pos_st"]) - (st + 50)) < 100  # the tolerance of the reference's own `eval`
        ok += int("+-"[g["strand"]] == strand and near)
    assert ok >= 0.85 * len(sigs), ok  # the reference's accuracy gate for DNA (test/test.sh:54-55)
    # same reads, different batch split / slot: identical records
    a = ctx.map_batch(sigs[:50], sc[:50], slot=1)
    b = ctx.map_batch(sigs[50:], sc[50:], slot=0)
    assert np.concatenate([a, b]).tobytes() == got.tobytes()
    ref = H.OracleRef([seq], lm, k, 0, 250)
    for i in (0, 7, 101):
        o = H.orc_map(ref, sigs[i], 8192.0, 10.0, 1402.882, 0, 250, 50)
        assert_hit_equal(got[i], o, ("c4", i), 0, 250, 50)
    ref.close()
    ctx.close()


@pytest.mark.parametrize("rna", [False, True])
def test_event_tables_at_tile_boundaries(rna):
    """signal lengths around the event kernel's tile (256), lag (32) and window (2w) boundaries"""
    rng = np.random.default_rng(21 + rna)
    k = 5 if rna else 6
    ctx = capi.Context(model(k), k, flags=H.F_RNA if rna else 0)
    sc = synth.RNA_SCALING if rna else synth.DNA_SCALING
    lens = [1, 5, 11, 12, 13, 27, 28, 29, 31, 32, 33, 63, 64, 65, 223, 224, 225, 255, 256, 257, 287, 288, 289, 511, 512, 513,
            767, 768, 769, 1023, 1024, 1025, 1279, 1280, 1281, 4095, 4096, 4097]
    lv = rng.uniform(60, 130, size=3000).astype(np.float32)
    base = synth.simulate_read(lv, rng, sc, min_dwell=3 if not rna else 6)
    checked = 0
    for n in lens:
        sig = base[100:100 + n].copy()
        ev = H.orc_events(sig, sc["digitisation"], sc["offset"], sc["range"], rna)
        start, length, mean = ctx.event_table(sig, sc)
        if len(ev) == 0:  # no peak: the reference is undefined here; the GPU reports an empty table
            assert len(start) == 0, n
            continue
        assert np.array_equal(start, ev["start"]), n
        assert np.array_equal(bits(length), bits(ev["length"])), n
        assert np.array_equal(bits(mean), bits(ev["mean"])), n
        checked += 1
    assert checked >= 20
    ctx.close()


def test_event_prefix_sums_inexact_fallback():
    """When the fp64 prefix sums are not exact a parallel scan could round differently from the reference's
    sequential sums: the kernel must notice (TwoSum) and redo the tile in order.  A 1e-8 offset next to
    zero-valued samples makes x*x span ~90 bits."""
    rng = np.random.default_rng(33)
    k = 6
    seqs = [synth.random_sequence(3000, rng)]
    lv = rng.uniform(60, 130, size=800).astype(np.float32)
    sc = dict(digitisation=8192.0, range=1402.882, offset=1e-8, sampling_rate=4000.0)
    sig = synth.simulate_read(lv, rng, synth.DNA_SCALING)
    sig[::7] = 0
    ctx = capi.Context(model(k), k)
    ctx.set_ref(seqs)
    ev = H.orc_events(sig, sc["digitisation"], sc["offset"], sc["range"], False)
    start, length, mean = ctx.event_table(sig, sc)
    assert np.array_equal(start, ev["start"])
    assert np.array_equal(bits(mean), bits(ev["mean"]))
    got = ctx.map_batch([sig], [sc])[0]
    assert got["status"] & 8, "the exactness guard did not fire: the test input no longer exercises the fallback"
    ref = H.OracleRef(seqs, model(k), k, 0, 250)
    o = H.orc_map(ref, sig, sc["digitisation"], sc["offset"], sc["range"], 0, 250, 50)
    assert_hit_equal(got, o, "inexact", 0, 250, 50)
    ref.close()
    ctx.close()


def test_many_short_references_more_groups_than_lanes():
    """load-balance shape of BASELINE.json configs[4]: many transcripts of <= 375 columns are packed into more
    than 32 segment groups per read (the per-read merge then needs several groups per lane); --rna --invert"""
    k = 5
    lm = model(k)
    rng = np.random.default_rng(31)
    seqs = [synth.random_sequence(int(n), rng) for n in rng.integers(380, 1500, size=1300)]
    flags = H.F_RNA | H.F_INV
    sigs, truth = synth.simulate_reads(seqs, k, lm, 6, seed=8, rna=True, bases_per_read=420)
    sc = [synth.RNA_SCALING] * len(sigs)
    ctx = capi.Context(lm, k, flags=flags)
    ctx.set_ref(seqs)
    assert ctx.ref_columns > 32 * 8192  # ~64 groups of ~7.6 k columns (sfgpu.cu:layout_ref)
    got = ctx.map_batch(sigs, sc)
    ref = H.OracleRef(seqs, lm, k, flags, 250)
    for i, s in enumerate(sigs):
        o = H.orc_map(ref, s, sc[i]["digitisation"], sc[i]["offset"], sc[i]["range"], flags, 250, 50)
        assert_hit_equal(got[i], o, ("many", i), flags, 250, 50)
    ref.close()
    ctx.close()


@pytest.mark.parametrize("q", [10, 40, 90, 128, 150, 180, 200, 256, 270, 320, 340, 380, 410, 440, 470, 512, 600, 760, 1000, 1024])
def test_dtw_every_register_tile_height(q):
    """one case per template instantiation of the DTW / trace kernels (R = 1 .. 16, 20, 24, 32):
    subsequence and standard DTW, full-length and ragged queries, checkpoints on"""
    rng = np.random.default_rng(5000 + q)
    lens = [int(x) for x in rng.integers(1, 700, size=4)] + [2 * q + 37, 2500]
    fwd = _rand_arrays(rng, lens, 2)
    rev = _rand_arrays(rng, lens, 2)
    qlens = [q, q, max(1, q - 1), max(1, (2 * q) // 3), min(q, 25)]
    queries = _rand_arrays(rng, qlens, 2)
    oref = H.OracleEventRef(fwd, rev)
    ctx = capi.Context(model(5), 5, query_size=q, ck_min_cols=256, min_window=8)
    ctx.set_ref_events(fwd, rev)
    got = ctx.align_queries(queries)
    for i, x in enumerate(queries):
        o = oref.align(x, 0)
        g = got[i]
        assert (g["rid"], "+-"[g["strand"]]) == (o.rid, o.strand.decode()), (q, i)
        assert bits(g["score"]) == bits(o.score) and bits(g["score2"]) == bits(o.score2), (q, i)
        assert (g["pos_st"], g["pos_end"]) == (o.raw_pos_st, o.raw_pos_end), (q, i)
    ctx.close()
    oref.close()
    oref = H.OracleEventRef(fwd, None)
    ctx = capi.Context(model(5), 5, flags=H.F_RNA | H.F_DTW, query_size=q, ck_min_cols=256, min_window=8)
    ctx.set_ref_events(fwd, None)
    got = ctx.align_queries(queries)
    for i, x in enumerate(queries):
        o = oref.align(x[::-1].copy(), H.F_RNA | H.F_DTW)
        g = got[i]
        assert g["rid"] == o.rid, (q, i)
        assert bits(g["score"]) == bits(o.score) and bits(g["score2"]) == bits(o.score2), (q, i)
        assert (g["pos_st"], g["pos_end"]) == (o.raw_pos_st, o.raw_pos_end), (q, i)
    ctx.close()
    oref.close()


def test_query_size_limit_is_reported():
    with pytest.raises(capi.SfgpuError, match="1024"):
        capi.Context(model(5), 5, query_size=1025)


def test_auto_query_start_on_truncated_reads():
    """-p -1 edge cases: reads cut inside the adaptor, inside the poly-A tail, right after it (start found but
    fewer than q / fewer than 25 events left: too short / ignored), reads shorter than the 2000-sample window
    of the adaptor finder, plus transcripts shorter than 1.5*q in the reference"""
    c = CASES["rna_tail24_auto"]
    ids, sigs, sc = H.case_reads(c)
    k, q = c["k"], c["q"]
    rng = np.random.default_rng(17)
    seqs = [synth.random_sequence(int(n), rng) for n in (300, 5, 60, 900, 379, 380, 2000)]
    cut = []
    for i, s in enumerate(sigs[:12]):
        for n in (1500, 2001, 3500, 5200, 6500, 7600, 9000, 12000):
            if n < len(s):
                cut.append(s[:n])
    scs = [synth.RNA_SCALING] * len(cut)
    ctx = capi.Context(model(k), k, flags=H.F_RNA, query_size=q, prefix_size=-1)
    ctx.set_ref(seqs)
    got = ctx.map_batch(cut, scs)
    ref = H.OracleRef(seqs, model(k), k, H.F_RNA, q)
    seen = set()
    for i, s in enumerate(cut):
        o = H.orc_map(ref, s, scs[i]["digitisation"], scs[i]["offset"], scs[i]["range"], H.F_RNA, q, -1)
        assert_hit_equal(got[i], o, ("cut", i, len(s)), H.F_RNA, q, -1)
        seen.add((bool(o.mapped), o.status))
    assert len(seen) >= 3, seen  # mapped / too short / ignored / prefix-fail combinations all occurred
    ref.close()
    ctx.close()


@pytest.mark.parametrize("seed", range(24))
def test_random_configurations_match_oracle(seed):
    """seeded fuzz over the option space: chemistry, flag combinations the CLI accepts, q, p, contig counts and
    lengths (down to a single k-mer), read lengths from 'ignored' to several thousand events"""
    rng = np.random.default_rng(9000 + seed)
    rna = bool(rng.integers(0, 2))
    k = (9 if seed % 3 == 2 else 5) if rna else int(rng.choice([6, 9]))  # RNA: r9 5-mers and RNA004 9-mers
    flags = 0
    if rna:
        flags = H.F_RNA
        flags |= int(rng.choice([0, H.F_DTW, H.F_INV, H.F_REF, H.F_DTW | H.F_REF, H.F_INV | H.F_REF]))
    if rng.integers(0, 3) == 0 and not (flags & H.F_INV and False):
        flags |= H.F_END
    q = int(rng.choice([25, 60, 97, 130, 250, 250, 333, 420]))
    p = int(rng.choice([0, 10, 50, 50, 120]))
    n_contig = int(rng.integers(1, 7))
    seqs = [synth.random_sequence(int(n), rng) for n in rng.integers(k, 4000, size=n_contig)]
    seqs[0] = synth.random_sequence(int(rng.integers(600, 5000)), rng)
    lm = model(k)
    sigs = []
    scs = []
    for r in range(8):
        nb = int(rng.choice([30, 80, 150, 300, 450, 700]))
        s, _ = synth.simulate_reads([seqs[0]], k, lm, 1, seed=int(rng.integers(1 << 30)), rna=rna, bases_per_read=nb,
                                    min_samples=600)
        sigs.append(s[0])
        base = synth.RNA_SCALING if rna else synth.DNA_SCALING
        scs.append(dict(base, offset=float(base["offset"] + rng.integers(-20, 20)), range=float(base["range"] * rng.uniform(0.9, 1.1))))
    ctx = capi.Context(lm, k, flags=flags, query_size=q, prefix_size=p)
    ctx.set_ref(seqs)
    got = ctx.map_batch(sigs, scs)
    ref = H.OracleRef(seqs, lm, k, flags, q)
    for i, s in enumerate(sigs):
        o = H.orc_map(ref, s, scs[i]["digitisation"], scs[i]["offset"], scs[i]["range"], flags, q, p)
        if o.mapped and (np.isnan(o.score) or np.isnan(o.score2)):
            continue  # a degenerate reference (single-column contig z-scores to NaN): undefined in the reference
        assert_hit_equal(got[i], o, (seed, i, flags, q, p), flags, q, p)
    ref.close()
    ctx.close()


def test_abi_misuse_is_reported_not_crashed():
    """error convention of include/sfgpu.h: negative code + text, no abort, context still usable"""
    import ctypes as C
    L = capi.lib()
    k = 6
    rng = np.random.default_rng(2)
    seqs = [synth.random_sequence(2000, rng)]
    sigs, _ = synth.simulate_reads(seqs, k, model(k), 3, seed=3, bases_per_read=400)
    sc = [synth.DNA_SCALING] * 3
    ctx = capi.Context(model(k), k)
    with pytest.raises(capi.SfgpuError, match="before sfgpu_set_ref"):
        ctx.submit(0, *ctx.pack(sigs, sc))
    ctx.set_ref(seqs)
    res = np.zeros(4, dtype=capi.RESULT_DTYPE)
    assert L.sfgpu_collect(ctx._h, 0, res.ctypes.data_as(C.c_void_p)) == -4  # nothing submitted
    assert L.sfgpu_collect(ctx._h, 7, res.ctypes.data_as(C.c_void_p)) == -2  # bad slot
    assert L.sfgpu_resubmit(ctx._h, 1) == -4
    buf = np.zeros(16, np.float32)
    assert L.sfgpu_ref_events(ctx._h, 3, 0, buf.ctypes.data_as(C.c_void_p), 16) == -2  # no such contig
    assert L.sfgpu_ref_events(ctx._h, 0, 0, buf.ctypes.data_as(C.c_void_p), 16) == -2  # buffer too small
    with pytest.raises(capi.SfgpuError, match="contig 0 has 3 bases"):
        ctx.set_ref([b"ACG"])
    ctx.set_ref(seqs)  # still usable after the failed call
    good = ctx.map_batch(sigs, sc)
    assert (good["qlen"] == 250).all()
    with pytest.raises(capi.SfgpuError, match="SFGPU_SAM"):
        ctx.collect_paths(0, good)
    qbuf = np.zeros(250, np.float32)
    qlen = np.array([300], np.int32)  # longer than query_size
    assert L.sfgpu_submit_queries(ctx._h, 0, 1, qbuf.ctypes.data_as(C.c_void_p), qlen.ctypes.data_as(C.c_void_p)) == -2
    assert b"qlen" in L.sfgpu_strerror(ctx._h)
    ctx.close()
    sam = capi.Context(model(k), k, flags=capi.SFGPU_SAM)
    sam.set_ref(seqs)
    r = sam.map_batch(sigs, sc)
    off = np.zeros(4, dtype=np.int64)  # no room for any move
    mv = np.zeros(8, np.uint8); nm = np.zeros(3, np.int32)
    es = np.zeros(3 * 250, np.uint64); el = np.zeros(3 * 250, np.float32)
    rc = L.sfgpu_collect_paths(sam._h, 0, off.ctypes.data_as(C.c_void_p), mv.ctypes.data_as(C.c_void_p),
                               nm.ctypes.data_as(C.c_void_p), es.ctypes.data_as(C.c_void_p), el.ctypes.data_as(C.c_void_p))
    assert rc == -2 and b"move buffer" in L.sfgpu_strerror(sam._h)
    paths, _, _ = sam.collect_paths(0, r)  # and the proper call still works afterwards
    assert all(p is not None for p in paths)
    sam.close()


def test_slow_and_fast_reads_in_one_batch():
    """reads translocating five times slower than the others (p+q events spread over > 13 k samples, many tiles
    of the event kernel) mixed with ordinary reads in one batch"""
    k = 6
    lm = model(k)
    rng = np.random.default_rng(41)
    seqs = [synth.random_sequence(5000, rng)]
    ranks_ = synth.kmer_ranks(seqs[0], k)
    sigs = []
    for i in range(10):
        st = int(rng.integers(0, len(ranks_) - 500))
        lv = lm[ranks_[st:st + 450]]
        slow = i % 2 == 0
        sigs.append(synth.simulate_read(lv, rng, synth.DNA_SCALING, mean_extra_dwell=45.0 if slow else 8.0, min_dwell=12 if slow else 2))
    sc = [synth.DNA_SCALING] * len(sigs)
    assert max(len(s) for s in sigs) > 3 * 4352
    ref = H.OracleRef(seqs, lm, k, 0, 250)
    ctx = capi.Context(lm, k)
    ctx.set_ref(seqs)
    got = ctx.map_batch(sigs, sc)
    for i, s in enumerate(sigs):
        o = H.orc_map(ref, s, 8192.0, 10.0, 1402.882, 0, 250, 50)
        assert_hit_equal(got[i], o, ("slow", i), 0, 250, 50)
    ref.close()
    ctx.close()


# the pair kernel holds R2 = ceil(q / 16) rows per lane and is instantiated for every register RQ = (q - 1) % R2 the
# last query row can sit in.  Here: every RQ of R2 = 16 (240 < q <= 256), and the first, a middle and the last RQ of
# R2 = 5 .. 15 (64 < q <= 240); test_pair_layout_every_query_size below visits all the others.
_PAIR_Q = list(range(241, 257)) + [q for r2 in range(5, 16) for q in (16 * (r2 - 1) + 1, 16 * (r2 - 1) + 8, 16 * r2)]


@pytest.mark.parametrize("q,std", [(q, False) for q in _PAIR_Q] + [(q, True) for q in (197, 250, 256)])
def test_paired_and_unpaired_layouts_agree(q, std):
    """128 < q <= 256: full-length reads run two per warp (sf_dtw_pair_kernel), the others one per warp.  Both
    layouts must give the oracle's answer and identical bytes; odd and even counts of full-length reads,
    ragged reads mixed in, checkpoints on (restart from half-warp checkpoints) and off"""
    rng = np.random.default_rng(5 * q + std)
    lens = [int(x) for x in rng.integers(1, 700, size=12)] + [5000, 1, q, q + 1, 2 * q + 3]
    fwd = _rand_arrays(rng, lens, 2)
    rev = None if std else _rand_arrays(rng, lens, 2)
    flags = (H.F_RNA | H.F_DTW) if std else 0
    oref = H.OracleEventRef(fwd, rev)
    for n_full in (1, 2, 5, 8):
        qlens = [q] * n_full + [q - 1, 1, q // 2]
        order = rng.permutation(len(qlens))
        queries = [_rand_arrays(rng, [qlens[j]], 2)[0] for j in order]
        outs = []
        for nopair in (2, True):  # 2: pair whatever the query size (--dtw-std pairs only 192 < q <= 256 by default)
            for ck, win in ((0, 0), (128, 1)):
                ctx = capi.Context(model(5), 5, flags=flags, query_size=q, ck_min_cols=ck, min_window=win, no_pairing=nopair)
                ctx.set_ref_events(fwd, rev)
                outs.append(ctx.align_queries(queries).tobytes())
                got = ctx.align_queries(queries)
                assert ctx.timing(0).dtw_launches in ((1, 2) if nopair is True else (2, 4))  # + the redo passes when segments are split
                ctx.close()
        assert all(o == outs[0] for o in outs), (q, std, n_full)
        for i, x in enumerate(queries):
            o = oref.align(x[::-1].copy() if std else x, flags)
            g = got[i]
            tag = (q, std, n_full, i)
            assert g["rid"] == o.rid and "+-"[g["strand"]] == o.strand.decode(), tag
            assert bits(g["score"]) == bits(o.score) and bits(g["score2"]) == bits(o.score2), tag
            assert (g["pos_st"], g["pos_end"]) == (o.raw_pos_st, o.raw_pos_end), tag
    oref.close()


def test_pair_layout_every_query_size():
    """one small case per instantiation of sf_dtw_pair_kernel<R2, false, RQ> and sf_trace_pair_kernel<R2>: every
    64 < q <= 256, full-length reads (odd count) + ragged ones, checkpoints and split pieces on, against the oracle"""
    rng = np.random.default_rng(77)
    lens = [int(x) for x in rng.integers(1, 500, size=5)] + [3000, 700]
    fwd = _rand_arrays(rng, lens, 2)
    rev = _rand_arrays(rng, lens, 2)
    oref = H.OracleEventRef(fwd, rev)
    for q in range(65, 257):
        qlens = [q, q, q - 1, q, q // 3]
        queries = _rand_arrays(rng, qlens, 2 if q % 2 else 0)
        ctx = capi.Context(model(5), 5, query_size=q, ck_min_cols=128, min_window=1, piece_periods=2, warm_blocks=9)
        ctx.set_ref_events(fwd, rev)
        got = ctx.align_queries(queries)
        assert ctx.timing(0).dtw_launches >= 2, q  # the pair kernel ran
        ctx.close()
        for i, x in enumerate(queries):
            o = oref.align(x, 0)
            g = got[i]
            assert (g["rid"], "+-"[g["strand"]]) == (o.rid, o.strand.decode()), (q, i)
            assert bits(g["score"]) == bits(o.score) and bits(g["score2"]) == bits(o.score2), (q, i)
            assert (g["pos_st"], g["pos_end"]) == (o.raw_pos_st, o.raw_pos_end), (q, i)
    oref.close()


def test_sparse_checkpoints_knob(monkeypatch):
    """SFGPU_CHECKPOINTS_PER_READ = 32: 16 x fewer wavefront checkpoints (the start-coordinate pass then recomputes
    longer windows, the pieces of split segments get longer): the same bytes as the default spacing, and the oracle's
    answer on a sample"""
    k = 6
    lm = model(k)
    rng = np.random.default_rng(31)
    seq = synth.random_sequence(300_000, rng)
    sigs, truth = synth.simulate_reads([seq], k, lm, 96, seed=5, bases_per_read=450)
    sc = [synth.DNA_SCALING] * len(sigs)
    ctx = capi.Context(lm, k)
    ctx.set_ref([seq])
    want = ctx.map_batch(sigs, sc)
    dense = ctx.timing(0).tasks_per_read
    ctx.close()
    monkeypatch.setenv("SFGPU_CHECKPOINTS_PER_READ", "32")
    ctx = capi.Context(lm, k)
    ctx.set_ref([seq])
    got = ctx.map_batch(sigs, sc)
    assert ctx.timing(0).tasks_per_read <= dense
    ctx.close()
    assert got.tobytes() == want.tobytes()
    ref = H.OracleRef([seq], lm, k, 0, 250)
    for i in (0, 50):
        o = H.orc_map(ref, sigs[i], 8192.0, 10.0, 1402.882, 0, 250, 50)
        assert_hit_equal(got[i], o, ("sparse", i), 0, 250, 50)
    ref.close()


def test_several_long_contigs_share_the_checkpoint_budget():
    """the checkpoint period is set by the total length of the long segments (~512 checkpoints per read whatever
    the genome size): six 40 kb contigs, both strands, production settings, every read against the oracle"""
    k = 6
    lm = model(k)
    rng = np.random.default_rng(23)
    seqs = [synth.random_sequence(40_000 + 1000 * i, rng) for i in range(6)]
    sigs, truth = synth.simulate_reads(seqs, k, lm, 18, seed=9, bases_per_read=450)
    sc = [synth.DNA_SCALING] * len(sigs)
    ctx = capi.Context(lm, k)
    ctx.set_ref(seqs)
    got = ctx.map_batch(sigs, sc)
    ref = H.OracleRef(seqs, lm, k, 0, 250)
    hits = 0
    for i, s in enumerate(sigs):
        o = H.orc_map(ref, s, 8192.0, 10.0, 1402.882, 0, 250, 50)
        assert_hit_equal(got[i], o, ("contigs", i), 0, 250, 50)
        hits += int(got[i]["rid"] == truth[i][0])
    assert hits >= 0.85 * len(sigs)
    ref.close()
    ctx.close()
