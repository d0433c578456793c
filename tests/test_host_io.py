"""CPU tests of the host-side readers (sigfish_b200/host): SLOW5 / BLOW5 records, FASTA, k-mer model.
They run without a GPU: libsfhost.so only needs libsfgpu.so to load."""
import ctypes as C
import gzip
import os
import shutil

import numpy as np
import pytest

import helpers as H
from sigfish_b200 import build as B
from sigfish_b200 import synth


class SfRec(C.Structure):
    _fields_ = [("read_id", C.c_char_p), ("digitisation", C.c_double), ("offset", C.c_double), ("range", C.c_double),
                ("sampling_rate", C.c_double), ("len_raw_signal", C.c_uint64), ("raw_signal", C.POINTER(C.c_int16)),
                ("cap_signal", C.c_size_t)]


class SfFasta(C.Structure):
    _fields_ = [("num_ref", C.c_int32), ("names", C.POINTER(C.c_char_p)), ("bases", C.POINTER(C.c_char)),
                ("off", C.POINTER(C.c_int64))]


@pytest.fixture(scope="module")
def host():
    B.build_all()
    L = C.CDLL(B.LIB_HOST)
    L.sf_s5_open.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
    L.sf_s5_open.restype = C.c_void_p
    L.sf_s5_close.argtypes = [C.c_void_p]
    L.sf_s5_hdr_get.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32]
    L.sf_s5_hdr_get.restype = C.c_char_p
    L.sf_s5_get_next_mem.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.sf_s5_get_next_mem.restype = C.c_int64
    L.sf_s5_parse.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(SfRec), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.sf_fasta_read.argtypes = [C.c_char_p, C.POINTER(SfFasta), C.c_char_p, C.c_size_t]
    L.sf_model_read.argtypes = [C.c_char_p, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_uint32), C.c_char_p, C.c_size_t]
    return L


def read_all(L, path):
    err = C.create_string_buffer(512)
    f = L.sf_s5_open(path.encode(), err, 512)
    assert f, err.value
    out = []
    mem, cap = C.c_void_p(), C.c_size_t(0)
    scratch, scap = C.c_void_p(), C.c_size_t(0)
    rec = SfRec()
    total = 0
    while True:
        n = L.sf_s5_get_next_mem(f, C.byref(mem), C.byref(cap))
        assert n >= 0
        if n == 0:
            break
        total += n
        assert L.sf_s5_parse(f, mem, n, C.byref(rec), C.byref(scratch), C.byref(scap)) == 0
        sig = np.ctypeslib.as_array(rec.raw_signal, shape=(rec.len_raw_signal,)).copy() if rec.len_raw_signal else np.zeros(0, np.int16)
        out.append((rec.read_id.decode(), rec.digitisation, rec.offset, rec.range, rec.sampling_rate, sig))
    kit = L.sf_s5_hdr_get(f, b"sequencing_kit", 0)
    exp = L.sf_s5_hdr_get(f, b"experiment_type", 0)
    assert L.sf_s5_hdr_get(f, b"no_such_attribute", 0) is None
    L.sf_s5_close(f)
    return out, kit, exp, total


@pytest.mark.parametrize("fmt", ["slow5", "blow5_zlib_svb", "blow5_zlib_raw", "blow5_none_svb", "blow5_none_raw"])
@pytest.mark.parametrize("name,rna", [("sp1_dna", False), ("sequin_rna", True)])
def test_slow5_blow5_round_trip(host, tmp_path, fmt, name, rna):
    ids, sigs, sc = H.load_reads_npz(os.path.join(H.GOLDEN, name + ".npz"))
    sigs = list(sigs) + [np.zeros(0, np.int16), np.array([-32768, 32767, 0, -1, 1], np.int16)]
    ids = ids + ["empty", "extremes"]
    sc = sc + [sc[0], sc[0]]
    p = str(tmp_path / ("r." + ("slow5" if fmt == "slow5" else "blow5")))
    if fmt == "slow5":
        synth.write_slow5_ascii(p, ids, sigs, rna=rna, scalings=sc)
    else:
        synth.write_blow5(p, ids, sigs, rna=rna, scalings=sc, record_zlib="zlib" in fmt, signal_svb="svb" in fmt)
    got, kit, exp, total = read_all(host, p)
    assert exp == (b"rna" if rna else b"genomic_dna") and kit is not None
    assert [g[0] for g in got] == ids
    for g, s, c in zip(got, sigs, sc):
        assert (g[1], g[2], g[3], g[4]) == (c["digitisation"], c["offset"], c["range"], c["sampling_rate"])
        assert np.array_equal(g[5], s)
    assert total > 0


@pytest.mark.refbin
@pytest.mark.parametrize("name", ["sp1_dna", "sequin_rna"])
def test_reads_bundled_reference_blow5(host, name):
    """the reference's own BLOW5 test files (zlib records, svb-zd signal, auxiliary fields), decoded by
    our reader, against the signals slow5lib decoded (tests/golden/*.npz).  Build container only."""
    p = f"/root/reference/test/{name}.blow5"
    if not os.path.exists(p):
        pytest.skip("reference mount not present")
    ids, sigs, sc = H.load_reads_npz(os.path.join(H.GOLDEN, name + ".npz"))
    got, kit, exp, _ = read_all(host, p)
    assert [g[0] for g in got] == ids
    for g, s, c in zip(got, sigs, sc):
        assert (g[1], g[2], g[3]) == (c["digitisation"], c["offset"], c["range"])
        assert np.array_equal(g[5], s)


def test_fasta_reader_plain_and_gzip(host, tmp_path):
    names, seqs = H.read_fasta(os.path.join(H.GOLDEN, "synth_multi.fa.gz"))
    plain = str(tmp_path / "m.fa")
    with gzip.open(os.path.join(H.GOLDEN, "synth_multi.fa.gz"), "rb") as fi, open(plain, "wb") as fo:
        shutil.copyfileobj(fi, fo)
    for p in (plain, os.path.join(H.GOLDEN, "synth_multi.fa.gz")):
        fa = SfFasta()
        err = C.create_string_buffer(512)
        assert host.sf_fasta_read(p.encode(), C.byref(fa), err, 512) == 0, err.value
        assert fa.num_ref == len(names)
        for i in range(fa.num_ref):
            assert fa.names[i].decode() == names[i]
            assert C.string_at(C.addressof(fa.bases.contents) + fa.off[i], fa.off[i + 1] - fa.off[i]) == seqs[i]
    # header with a description, Windows line ends, blank lines
    odd = str(tmp_path / "odd.fa")
    open(odd, "wb").write(b">c1 some description\r\nACGT\r\n\r\nacgtn\r\n>c2\tx\nTT\n")
    fa = SfFasta()
    err = C.create_string_buffer(512)
    assert host.sf_fasta_read(odd.encode(), C.byref(fa), err, 512) == 0
    assert [fa.names[i] for i in range(2)] == [b"c1", b"c2"]
    assert C.string_at(C.addressof(fa.bases.contents), fa.off[2]) == b"ACGTacgtnTT"
    assert host.sf_fasta_read(str(tmp_path / "missing.fa").encode(), C.byref(fa), err, 512) != 0


def test_model_reader(host, tmp_path):
    for k in (5, 6):
        mean, stdv = synth.make_model(k)
        p = str(tmp_path / f"m{k}.txt")
        synth.write_model_file(p, k, mean, stdv)
        lm = C.POINTER(C.c_float)()
        kk = C.c_uint32()
        err = C.create_string_buffer(512)
        assert host.sf_model_read(p.encode(), C.byref(lm), C.byref(kk), err, 512) == 0, err.value
        assert kk.value == k
        got = np.ctypeslib.as_array(lm, shape=(4 ** k,))
        assert np.array_equal(got.view(np.uint32), mean.view(np.uint32))
    # truncated file
    bad = str(tmp_path / "bad.txt")
    open(bad, "w").write("#k\t5\nAAAAA\t1.0\t1.0\n")
    lm = C.POINTER(C.c_float)()
    kk = C.c_uint32()
    err = C.create_string_buffer(512)
    assert host.sf_model_read(bad.encode(), C.byref(lm), C.byref(kk), err, 512) != 0
    assert b"prematurely" in err.value


@pytest.mark.refbin
@pytest.mark.parametrize("zl,svb", [(True, True), (False, True), (True, False), (False, False)])
def test_reference_binary_reads_the_blow5_we_write(tmp_path, zl, svb):
    """synth.write_blow5 is what the GPU-side CLI tests and tools/cli_e2e.py feed to BOTH binaries: slow5lib
    must accept it and the reference must print its golden PAF from it.  Build container only."""
    if not H.have_ref_bin():
        pytest.skip("oracle/_ref not built")
    ids, sigs, sc = H.load_reads_npz(os.path.join(H.GOLDEN, "sp1_dna.npz"))
    p = str(tmp_path / "r.blow5")
    synth.write_blow5(p, ids, sigs, scalings=sc, record_zlib=zl, signal_svb=svb)
    fa = str(tmp_path / "ref.fa")
    with gzip.open(os.path.join(H.GOLDEN, "nCoV-2019.fa.gz"), "rb") as fi, open(fa, "wb") as fo:
        shutil.copyfileobj(fi, fo)
    mean, stdv = synth.make_model(6)
    mf = str(tmp_path / "m.txt")
    synth.write_model_file(mf, 6, mean, stdv)
    assert H.run_ref(fa, p, mf) == open(os.path.join(H.GOLDEN, "paf", "dna_sp1_default.paf")).read()


# ---- sfinflate.c: the record decoder must agree with zlib byte for byte ----

def _inflater(host):
    host.sf_zlib_inflate.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    host.sf_zlib_inflate.restype = C.c_int
    state = C.create_string_buffer(1 << 16)  # sizeof(sf_inflater) < 32 KB, zero-initialised

    def run(data, cap):
        out = C.create_string_buffer(max(cap, 1))
        n = C.c_size_t(0)
        rc = host.sf_zlib_inflate(state, data, len(data), out, cap, C.byref(n))
        return rc, (out.raw[:n.value] if rc == 0 else b"")
    return run


def _payload(rng, kind, n):
    if kind == 0:
        return rng.integers(0, 256, n, dtype=np.uint8).tobytes()          # incompressible: stored / near-flat codes
    if kind == 1:
        return rng.integers(0, 4, n, dtype=np.uint8).tobytes()            # short codes
    if kind == 2:
        return bytes(n)                                                   # one long run: maximal matches, distance 1
    if kind == 3:
        return (np.cumsum(rng.integers(-3, 4, n)) & 255).astype(np.uint8).tobytes()
    if kind == 4:
        base = rng.integers(0, 256, max(1, n // 7), dtype=np.uint8).tobytes()
        return (base * 8)[:n]                                             # long-distance matches
    p = 1.0 / (np.arange(256) + 1.0) ** 2.2                               # long tail: codes beyond the first-level table
    return rng.choice(np.arange(256, dtype=np.uint8), n, p=p / p.sum()).tobytes()


@pytest.mark.parametrize("seed", range(6))
def test_inflate_matches_zlib_on_all_block_types(host, seed):
    import zlib
    run = _inflater(host)
    rng = np.random.default_rng(seed)
    for it in range(250):
        n = int(rng.choice([0, 1, 2, 7, 64, 1000, 9000, 70000, 200000]))
        raw = _payload(rng, int(rng.integers(0, 6)), n)
        strat = int(rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED]))
        co = zlib.compressobj(int(rng.integers(0, 10)), zlib.DEFLATED, int(rng.integers(9, 16)), 8, strat)
        comp = co.compress(raw[:n // 2]) + (co.flush(zlib.Z_FULL_FLUSH) if rng.random() < .3 else b"") + \
            co.compress(raw[n // 2:]) + co.flush()
        rc, got = run(comp, n + int(rng.integers(0, 40)))
        assert rc == 0 and got == raw, (seed, it, n, strat)
        if n > 0:  # output buffer too small is reported, never overrun
            assert run(comp, int(rng.integers(0, n)))[0] == 1
        # corrupted and truncated streams: rejected, or (if the damage is harmless) the same bytes zlib gives
        bad = bytearray(comp)
        for _ in range(int(rng.integers(1, 4))):
            bad[int(rng.integers(0, len(bad)))] ^= 1 << int(rng.integers(0, 8))
        if rng.random() < .3:
            bad = bad[:int(rng.integers(0, len(bad) + 1))]
        rc, got = run(bytes(bad), n + 64)
        assert rc in (0, 1, -1)
        if rc == 0:
            assert got == zlib.decompress(bytes(bad))


@pytest.mark.parametrize("seed", range(4))
def test_inflate_prefix_agrees_with_the_full_decoder(host, seed):
    """sf_zlib_inflate_prefix (table-free canonical decoder behind sf_s5_parse_head): whenever it answers, its bytes
    are the beginning of what zlib gives and their number is min(cap, 2 + u16 + tail); it declines stored / fixed
    first blocks and prefixes that do not lie inside the first block; damaged streams never make it overrun, and
    whatever it returns for them is what the full decoder returns for the same bytes"""
    import struct
    import zlib
    host.sf_zlib_inflate_prefix.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t,
                                            C.POINTER(C.c_size_t)]
    host.sf_zlib_inflate_prefix.restype = C.c_int
    state = C.create_string_buffer(1 << 16)
    full = _inflater(host)
    rng = np.random.default_rng(100 + seed)

    def prefix(data, cap, tail):
        out = C.create_string_buffer(cap + 64)
        out.raw = b"\xa5" * (cap + 64)
        n = C.c_size_t(0)
        rc = host.sf_zlib_inflate_prefix(state, data, len(data), out, cap, tail, C.byref(n))
        assert out.raw[cap:] == b"\xa5" * 64  # never writes past cap
        return rc, out.raw[:n.value]

    answered = declined = 0
    for it in range(400):
        idl = int(rng.choice([0, 1, 8, 36, 36, 36, 255, 300, 5000]))
        rid = bytes(rng.integers(48, 123, idl, dtype=np.uint8)) if rng.random() < .8 else b"a" * idl  # runs: matches
        n = int(rng.choice([0, 1, 3, 50, 600, 9000, 40000]))
        raw = struct.pack("<H", idl) + rid + struct.pack("<I4dQ", 0, 8192.0, float(rng.integers(-300, 300)), 1437.98, 4000.0, n) + \
            _payload(rng, int(rng.integers(0, 6)), n)
        strat = int(rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY,
                                zlib.Z_RLE, zlib.Z_FIXED]))
        co = zlib.compressobj(int(rng.integers(0, 10)), zlib.DEFLATED, int(rng.integers(9, 16)), 8, strat)
        cut = int(rng.integers(0, len(raw) + 1)) if rng.random() < .2 else len(raw)
        comp = co.compress(raw[:cut]) + (co.flush(zlib.Z_FULL_FLUSH) if cut < len(raw) else b"") + co.compress(raw[cut:]) + co.flush()
        cap = int(rng.choice([2, 3, 60, 384, 384, 384, 1000]))
        tail = int(rng.choice([0, 48, 48, 48, 100]))
        rc, got = prefix(comp, cap, tail)
        assert rc in (0, 1)
        if rc == 1:
            answered += 1
            assert len(got) == min(cap, 2 + idl + tail) and got == raw[:len(got)], (seed, it)
        else:
            declined += 1
        # damage: flipped bits, truncation
        bad = bytearray(comp)
        for _ in range(int(rng.integers(1, 4))):
            bad[int(rng.integers(0, len(bad)))] ^= 1 << int(rng.integers(0, 8))
        if rng.random() < .4:
            bad = bad[:int(rng.integers(0, len(bad) + 1))]
        rc, got = prefix(bytes(bad), cap, tail)
        if rc == 1:
            # the same bytes the table decoder produces when it is stopped at that many output bytes
            out = C.create_string_buffer(len(got) + 1)
            m = C.c_size_t(0)
            rc2 = host.sf_zlib_inflate(C.create_string_buffer(1 << 16), bytes(bad), len(bad), out, len(got), C.byref(m))
            if rc2 == 1:
                assert out.raw[:m.value] == got[:m.value], (seed, it)
            elif rc2 == 0:
                assert out.raw[:m.value] == got
    assert answered > 100 and declined > 20, (answered, declined)


def test_blow5_records_decode_identically_with_every_codec_combination(host, tmp_path):
    """zlib on / off x svb-zd on / off, signals with the extreme deltas (4-byte svb codes) and lengths that
    are not a multiple of four (the tail of the fast svb loop)"""
    rng = np.random.default_rng(11)
    sigs = [rng.integers(-32768, 32768, size=int(n), dtype=np.int16) for n in (0, 1, 2, 3, 4, 5, 17, 1023, 4096, 9001)]
    sigs += [np.cumsum(rng.integers(-40, 41, size=6000)).astype(np.int16), np.full(3000, -32768, dtype=np.int16)]
    ids = [f"r{i}" for i in range(len(sigs))]
    for zl in (True, False):
        for svb in (True, False):
            p = str(tmp_path / f"c_{int(zl)}{int(svb)}.blow5")
            synth.write_blow5(p, ids, sigs, record_zlib=zl, signal_svb=svb)
            recs = read_all(host, p)[0]
            assert [r[0] for r in recs] == ids
            for r, s in zip(recs, sigs):
                assert np.array_equal(r[5], s)


@pytest.mark.parametrize("name", ["sp1_dna", "sequin_rna"])
def test_real_blow5_of_the_reference_decodes_to_the_golden_signals(host, name):
    """the reference's own test files (written by slow5tools: zlib records, svb-zd signal) through s5read.c and
    sfinflate.c; the expected signals are the committed golden arrays.  Needs the reference mount."""
    path = f"/root/reference/test/{name}.blow5"
    if not os.path.exists(path):
        pytest.skip("reference mount not present")
    import struct
    import zlib
    recs = read_all(host, path)[0]
    ids, sigs, _ = H.load_reads_npz(os.path.join(H.GOLDEN, name + ".npz"))
    assert [r[0] for r in recs] == list(ids)
    for r, s in zip(recs, sigs):
        assert np.array_equal(r[5], s)
    # and the fast decoder handles these streams itself (no zlib fallback involved)
    run = _inflater(host)
    b = open(path, "rb").read()
    o = 68 + struct.unpack_from("<I", b, 64)[0]
    n = 0
    while b[o:o + 5] != b"5WOLB":
        sz = struct.unpack_from("<Q", b, o)[0]
        rec = b[o + 8:o + 8 + sz]
        o += 8 + sz
        want = zlib.decompress(rec)
        rc, got = run(rec, len(want) + 8)
        assert rc == 0 and got == want
        n += 1
    assert n == len(ids)


def _blow5_with_record(path, rec_bytes, record_zlib, signal_svb, hdr_extra=b"", num_groups=1):
    """a BLOW5 v0.2.0 file holding one hand-made record"""
    import struct
    import zlib
    hdr = (b"@experiment_type\tgenomic_dna\n@sequencing_kit\tsqk-lsk109\n" + hdr_extra +
           b"#char*\tuint32_t\tdouble\tdouble\tdouble\tdouble\tuint64_t\tint16_t*\n"
           b"#read_id\tread_group\tdigitisation\toffset\trange\tsampling_rate\tlen_raw_signal\traw_signal\n")
    head = b"BLOW5\x01" + bytes([0, 2, 0]) + bytes([1 if record_zlib else 0]) + struct.pack("<I", num_groups) + \
        bytes([1 if signal_svb else 0])
    if record_zlib and rec_bytes is not None:
        rec_bytes = zlib.compress(rec_bytes)
    with open(path, "wb") as f:
        f.write(head + b"\x00" * (64 - len(head)) + struct.pack("<I", len(hdr)) + hdr)
        if rec_bytes is not None:
            f.write(struct.pack("<Q", len(rec_bytes)) + rec_bytes)
        f.write(b"5WOLB")


def _parse_first(L, path):
    err = C.create_string_buffer(512)
    f = L.sf_s5_open(path.encode(), err, 512)
    assert f, err.value
    mem, cap = C.c_void_p(), C.c_size_t(0)
    scratch, scap = C.c_void_p(), C.c_size_t(0)
    rec = SfRec()
    n = L.sf_s5_get_next_mem(f, C.byref(mem), C.byref(cap))
    assert n > 0
    rc = L.sf_s5_parse(f, mem, n, C.byref(rec), C.byref(scratch), C.byref(scap))
    L.sf_s5_close(f)
    return rc, rec


@pytest.mark.parametrize("record_zlib", [False, True])
@pytest.mark.parametrize("signal_svb", [False, True])
@pytest.mark.parametrize("length", [1 << 63, (1 << 64) - 1, (1 << 63) + 7, 1 << 40, 1 << 32])
def test_malformed_record_lengths_are_rejected(host, tmp_path, record_zlib, signal_svb, length):
    """a record whose len_raw_signal field wraps every size computation: must be refused, never
    reported as parsed with a wrapped buffer"""
    import struct
    rid = b"evil"
    body = struct.pack("<H", len(rid)) + rid + struct.pack("<I", 0) + struct.pack("<dddd", 8192.0, 10.0, 1402.0, 4000.0) + \
        struct.pack("<Q", length) + b"\x01\x02\x03\x04\x05\x06\x07\x08"
    p = str(tmp_path / "evil.blow5")
    _blow5_with_record(p, body, record_zlib, signal_svb)
    rc, _ = _parse_first(host, p)
    assert rc != 0


def test_malformed_svb_count_and_truncated_zlib_are_rejected(host, tmp_path):
    import struct
    import zlib
    rid = b"evil"
    pre = struct.pack("<H", len(rid)) + rid + struct.pack("<I", 0) + struct.pack("<dddd", 8192.0, 10.0, 1402.0, 4000.0)
    # svb-zd stream announcing 2^32-1 samples in 12 bytes
    svb = struct.pack("<I", 0xffffffff) + b"\x00" * 8
    p = str(tmp_path / "svb.blow5")
    _blow5_with_record(p, pre + struct.pack("<Q", len(svb)) + svb, False, True)
    assert _parse_first(host, p)[0] != 0
    # a zlib stream cut short: must fail instead of doubling the output buffer for ever
    good = pre + struct.pack("<Q", 64) + np.arange(64, dtype=np.int16).tobytes()[:128]
    z = zlib.compress(pre + struct.pack("<Q", 4000) + np.arange(4000, dtype=np.int16).tobytes())
    with open(str(tmp_path / "cut.blow5"), "wb") as f:
        head = b"BLOW5\x01" + bytes([0, 2, 0, 1]) + struct.pack("<I", 1) + bytes([0])
        hdr = b"@experiment_type\tgenomic_dna\n#read_id\n"
        f.write(head + b"\x00" * (64 - len(head)) + struct.pack("<I", len(hdr)) + hdr)
        cut = z[:len(z) // 2]
        f.write(struct.pack("<Q", len(cut)) + cut + b"5WOLB")
    assert _parse_first(host, str(tmp_path / "cut.blow5"))[0] != 0
    del good
    # ASCII: a length field far larger than the signal column
    s5 = str(tmp_path / "evil.slow5")
    with open(s5, "w") as f:
        f.write("#slow5_version\t0.2.0\n#num_read_groups\t1\n@experiment_type\tgenomic_dna\n"
                "#char*\tuint32_t\tdouble\tdouble\tdouble\tdouble\tuint64_t\tint16_t*\n"
                "#read_id\tread_group\tdigitisation\toffset\trange\tsampling_rate\tlen_raw_signal\traw_signal\n"
                "evil\t0\t8192\t10\t1402\t4000\t9223372036854775808\t1,2,3\n")
    assert _parse_first(host, s5)[0] != 0


def test_header_read_group_count_cannot_grow_after_attributes(host, tmp_path):
    """'#num_read_groups' after '@' rows (or disagreeing with the binary header) must not widen the rows that
    sf_s5_hdr_get / sf_s5_close index"""
    import struct
    rid = b"r0"
    body = struct.pack("<H", len(rid)) + rid + struct.pack("<I", 0) + struct.pack("<dddd", 8192.0, 10.0, 1402.0, 4000.0) + \
        struct.pack("<Q", 2) + np.array([5, 6], np.int16).tobytes()
    p = str(tmp_path / "g.blow5")
    _blow5_with_record(p, body, False, False, hdr_extra=b"#num_read_groups\t4000000\n@late\tx\n")
    err = C.create_string_buffer(512)
    f = host.sf_s5_open(p.encode(), err, 512)
    assert f, err.value
    assert host.sf_s5_hdr_get(f, b"sequencing_kit", 0) == b"sqk-lsk109"
    for g in (1, 2, 1000, 3999999):
        assert host.sf_s5_hdr_get(f, b"sequencing_kit", g) is None
        assert host.sf_s5_hdr_get(f, b"late", g) is None
    host.sf_s5_close(f)
    # the same in a text file
    s5 = str(tmp_path / "g.slow5")
    with open(s5, "w") as fo:
        fo.write("#slow5_version\t0.2.0\n#num_read_groups\t1\n@experiment_type\trna\n#num_read_groups\t70000\n@late\tx\n"
                 "#char*\tuint32_t\n#read_id\tread_group\n")
    f = host.sf_s5_open(s5.encode(), err, 512)
    assert f, err.value
    assert host.sf_s5_hdr_get(f, b"experiment_type", 0) == b"rna"
    assert host.sf_s5_hdr_get(f, b"experiment_type", 5) is None
    host.sf_s5_close(f)
    # an absurd count in the binary header is refused at open
    _blow5_with_record(p, body, False, False, num_groups=0x7fffffff)
    assert not host.sf_s5_open(p.encode(), err, 512)


@pytest.mark.parametrize("fmt", ["blow5_zlib_svb", "blow5_zlib_raw", "blow5_none_svb", "blow5_none_raw", "reference_file"])
def test_record_heads_agree_with_full_decode(host, tmp_path, fmt):
    """sf_s5_parse_head (what the host reads when the GPUs decode the records): id, scaling and sample count equal
    the full decoder's, and the signal field it points at really is the signal (checked by decoding it here)"""
    import struct
    import zlib
    host.sf_s5_parse_head.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(SfRec), C.POINTER(C.c_int32),
                                      C.POINTER(C.c_int64), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    host.sf_s5_get_next_view.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    host.sf_s5_get_next_view.restype = C.c_int64
    host.sf_s5_is_mapped.argtypes = [C.c_void_p]
    ids, sigs, sc = H.load_reads_npz(os.path.join(H.GOLDEN, "sp1_dna.npz"))
    if fmt == "reference_file":
        p = os.path.join(H.GOLDEN, "sp1_dna.blow5")
        zl, svb = True, True
    else:
        sigs = list(sigs) + [np.zeros(0, np.int16), np.array([-32768, 32767, 0, -1, 1], np.int16)]
        ids = ids + ["empty", "extremes"]
        sc = sc + [sc[0], sc[0]]
        p = str(tmp_path / "r.blow5")
        zl, svb = "zlib" in fmt, "svb" in fmt
        synth.write_blow5(p, ids, sigs, scalings=sc, record_zlib=zl, signal_svb=svb)
    err = C.create_string_buffer(512)
    f = host.sf_s5_open(p.encode(), err, 512)
    assert f and host.sf_s5_is_mapped(f) == 1
    scratch, scap = C.c_void_p(), C.c_size_t(0)
    for i in range(len(ids)):
        view = C.c_void_p()
        n = host.sf_s5_get_next_view(f, C.byref(view))
        assert n > 0
        rec, pos, nbytes = SfRec(), C.c_int32(), C.c_int64()
        assert host.sf_s5_parse_head(f, view, n, C.byref(rec), C.byref(pos), C.byref(nbytes), C.byref(scratch), C.byref(scap)) == 0
        assert rec.read_id.decode() == ids[i]
        assert (rec.digitisation, rec.offset, rec.range) == (sc[i]["digitisation"], sc[i]["offset"], sc[i]["range"])
        assert rec.len_raw_signal == len(sigs[i])
        raw = C.string_at(view, n)
        body = zlib.decompress(raw) if zl else raw
        field = body[pos.value:pos.value + nbytes.value]
        assert len(field) == nbytes.value
        if svb:
            assert struct.unpack("<I", field[:4])[0] == len(sigs[i])
            assert field == synth._svb_zd_encode(np.asarray(sigs[i], np.int16))
        else:
            assert field == np.asarray(sigs[i], np.int16).tobytes()
    assert host.sf_s5_get_next_view(f, C.byref(view)) == 0
    host.sf_s5_close(f)


@pytest.mark.parametrize("path", [os.path.join(H.GOLDEN, "sp1_dna.blow5"), "/root/reference/test/sp1_dna.blow5",
                                  "/root/reference/test/sequin_rna.blow5"])
def test_record_heads_of_slow5tools_files(host, path):
    """files written by slow5tools (the reference's own test data: zlib records with auxiliary fields, svb-zd signals):
    the head of every record -- read by the table-free prefix decoder -- holds the id, scaling and sample count the
    full decoder finds, and the signal field it points at decodes to the same samples"""
    if not os.path.exists(path):
        pytest.skip("reference mount not present")
    import struct
    import zlib
    full = read_all(host, path)[0]
    host.sf_zlib_inflate_prefix.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t,
                                            C.POINTER(C.c_size_t)]
    host.sf_zlib_inflate_prefix.restype = C.c_int
    err = C.create_string_buffer(512)
    f = host.sf_s5_open(path.encode(), err, 512)
    assert f and host.sf_s5_record_press(f) == 1 and host.sf_s5_signal_press(f) == 1
    scratch, scap = C.c_void_p(), C.c_size_t(0)
    state, out, got = C.create_string_buffer(1 << 16), C.create_string_buffer(512), C.c_size_t(0)
    answered = 0
    for want in full:
        view = C.c_void_p()
        n = host.sf_s5_get_next_view(f, C.byref(view))
        assert n > 0
        rec, pos, nbytes = SfRec(), C.c_int32(), C.c_int64()
        assert host.sf_s5_parse_head(f, view, n, C.byref(rec), C.byref(pos), C.byref(nbytes), C.byref(scratch), C.byref(scap)) == 0
        assert (rec.read_id.decode(), rec.digitisation, rec.offset, rec.range, rec.sampling_rate) == want[:5]
        assert rec.len_raw_signal == len(want[5])
        raw = C.string_at(view, n)
        body = zlib.decompress(raw)
        field = body[pos.value:pos.value + nbytes.value]
        assert struct.unpack("<I", field[:4])[0] == len(want[5])
        assert field == synth._svb_zd_encode(want[5])
        answered += host.sf_zlib_inflate_prefix(state, raw, len(raw), out, 384, 48, C.byref(got))
        assert out.raw[:got.value] == body[:got.value]
    assert answered == len(full)  # these streams start with a dynamic block that holds the whole head
    host.sf_s5_close(f)
