"""BLOW5 records decoded on the device (sfgpu_submit_records: zlib inflate + svb-zd in csrc/sf_blow5.cuh) against the
host decoders: the samples the device produced must equal the input signals bit for bit, and the mapping results must
be byte-identical to the same reads submitted as int16 samples."""
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

import helpers as H
from sigfish_b200 import build as B
from sigfish_b200 import capi, synth

pytestmark = pytest.mark.gpu


def make_records(ids, sigs, scs, record_zlib, signal_svb, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, aux=b""):
    """BLOW5 records (what follows the u64 size in the file) + the head fields sfgpu_submit_records takes"""
    recs, sig_pos, sig_bytes = [], [], []
    for rid, sig, sc in zip(ids, sigs, scs):
        sig = np.ascontiguousarray(sig, dtype=np.int16)
        body = synth._svb_zd_encode(sig) if signal_svb else sig.tobytes()
        n_field = len(body) if signal_svb else len(sig)
        rb = rid.encode()
        head = struct.pack("<H", len(rb)) + rb + struct.pack("<I", 0) + \
            struct.pack("<dddd", sc["digitisation"], sc["offset"], sc["range"], sc["sampling_rate"]) + struct.pack("<Q", n_field)
        rec = head + body + aux
        if record_zlib:
            co = zlib.compressobj(level, zlib.DEFLATED, 15, 8, strategy)
            rec = co.compress(rec) + co.flush()
        recs.append(rec)
        sig_pos.append(len(head))
        sig_bytes.append(len(body))
    return recs, sig_pos, sig_bytes


@pytest.fixture(scope="module")
def setup():
    k = 6
    mean, _ = synth.make_model(k)
    rng = np.random.default_rng(77)
    seqs = [synth.random_sequence(12000, rng)]
    sigs, _ = synth.simulate_reads(seqs, k, mean, 40, seed=78, bases_per_read=430)
    # odd cases: empty read, extremes, a constant run (long matches), a long read (several deflate blocks), tiny reads
    sigs += [np.zeros(0, np.int16), np.array([-32768, 32767, 0, -1, 1, 32767, -32768], np.int16), np.full(5000, 517, np.int16),
             np.concatenate([sigs[0]] * 30), np.array([5], np.int16), np.arange(-300, 300, dtype=np.int16)]
    ids = [f"read_{i:04d}_{'x' * (i % 7)}" for i in range(len(sigs))]
    scs = [synth.DNA_SCALING] * len(sigs)
    ctx = capi.Context(mean, k)
    ctx.set_ref(seqs)
    want = ctx.map_batch(sigs, scs).copy()
    yield ctx, ids, sigs, scs, want
    ctx.close()


@pytest.mark.parametrize("record_zlib,signal_svb", [(True, True), (True, False), (False, True), (False, False)])
@pytest.mark.parametrize("level,strategy,aux", [(6, zlib.Z_DEFAULT_STRATEGY, b""), (1, zlib.Z_DEFAULT_STRATEGY, b"\x01\x02aux fields follow the signal" * 3),
                                                 (9, zlib.Z_DEFAULT_STRATEGY, b""), (0, zlib.Z_DEFAULT_STRATEGY, b""),
                                                 (6, zlib.Z_FIXED, b"tail"), (6, zlib.Z_HUFFMAN_ONLY, b""), (6, zlib.Z_RLE, b"")])
def test_device_decoding_matches_host(setup, record_zlib, signal_svb, level, strategy, aux):
    ctx, ids, sigs, scs, want = setup
    if not record_zlib and (level, strategy) != (6, zlib.Z_DEFAULT_STRATEGY):
        pytest.skip("compression settings only matter for zlib records")
    recs, sig_pos, sig_bytes = make_records(ids, sigs, scs, record_zlib, signal_svb, level, strategy, aux)
    ctx.submit_records(1, recs, int(record_zlib), int(signal_svb), sig_pos, sig_bytes, [len(s) for s in sigs], scs)
    got = ctx.collect(1)
    for i, s in enumerate(sigs):
        dev = ctx.slot_signal(1, i, len(s))
        assert np.array_equal(dev, s), (i, len(s))
    assert got.tobytes() == want.tobytes()


def test_malformed_records_are_reported_not_mapped(setup):
    ctx, ids, sigs, scs, want = setup
    recs, sig_pos, sig_bytes = make_records(ids, sigs, scs, True, True)
    ns = [len(s) for s in sigs]
    for what in ("flip", "truncate", "not_zlib", "short_signal"):
        bad = list(recs)
        sp, sb = list(sig_pos), list(sig_bytes)
        if what == "flip":       # one bit of the compressed data: the Huffman stream derails or the checksum fails
            b = bytearray(bad[3]); b[len(b) // 2] ^= 0x10; bad[3] = bytes(b)
        elif what == "truncate":
            bad[5] = bad[5][:len(bad[5]) * 2 // 3]
        elif what == "not_zlib":
            bad[7] = b"\x00\x01" + bad[7][2:]
        else:                    # the head announces more signal bytes than the record inflates to
            sb[2] += 4096
        ctx.submit_records(1, bad, 1, 1, sp, sb, ns, scs)
        with pytest.raises(capi.SfgpuError, match="could not be decoded on the device"):
            ctx.collect(1)
    # the context is still usable
    ctx.submit_records(1, recs, 1, 1, sig_pos, sig_bytes, ns, scs)
    assert ctx.collect(1).tobytes() == want.tobytes()


def _cli(args):
    B.build_all()
    r = subprocess.run([B.CLI, "dtw"] + args + ["--gpus", "1"], capture_output=True, text=True)
    return r


def test_cli_on_the_reference_blow5_file_with_and_without_device_decoding(tmp_path):
    """test/sp1_dna.blow5 of the reference as slow5lib wrote it (zlib records, svb-zd signals, auxiliary fields),
    a copy of which is kept under tests/golden: the PAF must equal the golden one whichever side decodes"""
    import gzip, json, shutil
    c = json.load(open(os.path.join(H.GOLDEN, "cases.json")))["dna_sp1_default"]
    fa, mf = str(tmp_path / "ref.fa"), str(tmp_path / "model.txt")
    with gzip.open(os.path.join(H.GOLDEN, "nCoV-2019.fa.gz"), "rb") as fi, open(fa, "wb") as fo:
        shutil.copyfileobj(fi, fo)
    synth.write_model_file(mf, c["k"], *synth.make_model(c["k"]))
    want = open(os.path.join(H.GOLDEN, "paf", "dna_sp1_default.paf")).read()
    for mode in ("yes", "no"):
        r = _cli([fa, os.path.join(H.GOLDEN, "sp1_dna.blow5"), "--kmer-model", mf, "--device-decode=" + mode])
        assert r.returncode == 0, r.stderr[-1500:]
        assert r.stdout == want, mode


def test_cli_corrupt_record_fails_loudly(tmp_path):
    """a damaged record: the device reports it, the host decoder gets the batch, finds the same damage and stops"""
    k = 6
    mean, stdv = synth.make_model(k)
    rng = np.random.default_rng(3)
    seqs = [synth.random_sequence(5000, rng)]
    sigs, _ = synth.simulate_reads(seqs, k, mean, 12, seed=4, bases_per_read=400)
    fa, s5, mf = str(tmp_path / "ref.fa"), str(tmp_path / "reads.blow5"), str(tmp_path / "model.txt")
    synth.write_fasta(fa, ["c0"], seqs)
    synth.write_model_file(mf, k, mean, stdv)
    synth.write_blow5(s5, [f"r{i}" for i in range(12)], sigs)
    raw = bytearray(open(s5, "rb").read())
    raw[len(raw) // 2] ^= 0x04
    open(s5, "wb").write(bytes(raw))
    r = _cli([fa, s5, "--kmer-model", mf])
    assert r.returncode != 0
    assert "decoding this batch on the host" in r.stderr and "Error parsing the record" in r.stderr
