"""End-to-end drop-in test on the GPU: `sigfish-b200 dtw` (C host + libsfgpu) must print, byte for byte,
the PAF the unmodified reference binary printed for the same inputs (tests/golden/paf/*.paf)."""
import gzip
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

import helpers as H
from sigfish_b200 import build as B
from sigfish_b200 import capi, synth

pytestmark = pytest.mark.gpu

CASES = json.load(open(os.path.join(H.GOLDEN, "cases.json")))


@pytest.fixture(scope="module")
def cli():
    B.build_all()
    assert os.path.exists(B.CLI)
    return B.CLI


def _inputs(tmp, case, fmt):
    c = CASES[case]
    fa = os.path.join(tmp, "ref.fa")
    if c["fasta"] in H.GENERATED_FASTA:
        H.write_case_fasta(c, fa)
    else:
        with gzip.open(os.path.join(H.GOLDEN, c["fasta"] + ".fa.gz"), "rb") as fi, open(fa, "wb") as fo:
            shutil.copyfileobj(fi, fo)
    ids, sigs, sc = H.case_reads(c)
    rna = bool(c["flags"] & H.F_RNA)
    kit = H.case_kit(c)
    reads = os.path.join(tmp, "reads." + fmt)
    if fmt == "slow5":
        synth.write_slow5_ascii(reads, ids, sigs, rna=rna, kit=kit, scalings=sc)
    else:
        synth.write_blow5(reads, ids, sigs, rna=rna, kit=kit, scalings=sc)
    mean, stdv = synth.make_model(c["k"])
    mf = os.path.join(tmp, "model.txt")
    synth.write_model_file(mf, c["k"], mean, stdv)
    return c, fa, reads, mf


def _run(cli, c, fa, reads, mf, extra=()):
    extra = list(extra)
    if "--gpus" not in extra:  # the CLI takes every visible GPU by default; one is enough (and much faster to set up)
        extra += ["--gpus", "1"]
    cmd = [cli, "dtw", fa, reads, "--kmer-model", mf, "-q", str(c["q"]), "-p", str(c["p"])] + H.flags_to_cli(c["flags"]) + extra
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout, r.stderr


@pytest.mark.parametrize("case", sorted(CASES))
def test_cli_paf_is_byte_identical_to_reference(cli, tmp_path, case):
    fmt = "blow5" if sum(map(ord, case)) % 2 else "slow5"
    c, fa, reads, mf = _inputs(str(tmp_path), case, fmt)
    out, err = _run(cli, c, fa, reads, mf)
    want = open(os.path.join(H.GOLDEN, "paf", case + ".paf")).read()
    assert out == want
    assert "total entries" in err and "DTW time" in err


@pytest.mark.parametrize("case,extra", [("dna_synth48", ["-K", "7", "-t", "3"]), ("dna_synth48", ["-B", "30K"]),
                                        ("rna_synth32", ["-K", "32"]), ("dna_multi_contig", ["-K", "1"]),
                                        ("dna_synth48", ["--debug-break", "1", "-K", "10"])])
def test_cli_batching_does_not_change_output(cli, tmp_path, case, extra):
    c, fa, reads, mf = _inputs(str(tmp_path), case, "blow5")
    out, _ = _run(cli, c, fa, reads, mf, extra)
    want = open(os.path.join(H.GOLDEN, "paf", case + ".paf")).read()
    if "--debug-break" in extra:  # stops after N+1 batches (dtw_main.c:322-325)
        want = "".join(want.splitlines(keepends=True)[:20])
    assert out == want


def test_cli_read_sharding_over_gpus(cli, tmp_path):
    n = capi.lib().sfgpu_device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    c, fa, reads, mf = _inputs(str(tmp_path), "dna_synth48", "blow5")
    want = open(os.path.join(H.GOLDEN, "paf", "dna_synth48.paf")).read()
    for g in sorted({2, n}):
        out, err = _run(cli, c, fa, reads, mf, ["--gpus", str(g), "-K", "20"])
        assert out == want
        assert f"on {g} GPU(s)" in err


def test_cli_read_sharding_sam_and_auto_start(cli, tmp_path):
    """the per-shard path / window-event buffers of --sam and the -p -1 kernels with reads split over GPUs"""
    n = capi.lib().sfgpu_device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    strip = lambda t: "".join(l for l in t.splitlines(keepends=True) if not l.startswith("@PG"))
    for case in ("rna_tail24_auto", "dna_synth48"):
        c, fa, reads, mf = _inputs(str(tmp_path), case, "blow5")
        want = open(os.path.join(H.GOLDEN, "sam", case + ".sam")).read()
        out, _ = _run(cli, c, fa, reads, mf, ["--sam", "--gpus", "2", "-K", "9"])
        assert strip(out) == strip(want), case
        out, _ = _run(cli, c, fa, reads, mf, ["--gpus", "2"])
        assert out == open(os.path.join(H.GOLDEN, "paf", case + ".paf")).read(), case


def test_cli_errors_like_the_reference(cli, tmp_path):
    c, fa, reads, mf = _inputs(str(tmp_path), "dna_sp1_default", "slow5")
    for extra, msg in ((["--dtw-std"], "DTW is only available for RNA"), (["--invert"], "Inversion is only available for RNA"),
                       (["--full-ref"], "--full-ref is only available for RNA"), (["-p", "-1"], "DNA does not support auto query start"),
                       (["--pore", "r11"], "Pore model should be")):
        r = subprocess.run([cli, "dtw", fa, reads, "--kmer-model", mf, "--gpus", "1"] + extra, capture_output=True, text=True)
        assert r.returncode != 0 and msg in r.stderr, (extra, r.stderr)
    r = subprocess.run([cli, "dtw", fa, reads], capture_output=True, text=True)
    assert r.returncode != 0 and "--kmer-model" in r.stderr
    r = subprocess.run([cli, "dtw", fa], capture_output=True, text=True)
    assert r.returncode != 0 and "Usage: sigfish dtw" in r.stderr
    r = subprocess.run([cli, "dtw", fa, str(tmp_path / "missing.blow5"), "--kmer-model", mf], capture_output=True, text=True)
    assert r.returncode != 0 and "Error opening SLOW5 file" in r.stderr


SAM_CASES = sorted(f[:-4] for f in os.listdir(os.path.join(H.GOLDEN, "sam")))


@pytest.mark.parametrize("case", SAM_CASES)
def test_cli_sam_is_byte_identical_to_reference(cli, tmp_path, case):
    """--sam: @SQ header + one record per read with the si:Z / ss:Z tags (needs the winner's full warping path)"""
    c, fa, reads, mf = _inputs(str(tmp_path), case, "blow5")
    out, _ = _run(cli, c, fa, reads, mf, ["--sam", "-K", "5"])
    want = open(os.path.join(H.GOLDEN, "sam", case + ".sam")).read()
    strip = lambda t: "".join(l for l in t.splitlines(keepends=True) if not l.startswith("@PG"))
    assert strip(out) == strip(want)
    assert out.count("@PG\tID:sigfish") == 1


def test_cli_large_batches_use_the_threaded_epilogue(cli, tmp_path):
    """batches of >= 1024 reads run the per-read epilogue (flip, offsets, MAPQ, PAF / SAM text) on the -t worker threads
    in chunks of 512 reads: the output must be, byte for byte, what small batches (epilogue on the main thread) give,
    for PAF and for SAM"""
    k = 6
    mean, stdv = synth.make_model(k)
    seq = synth.random_sequence(12_000, np.random.default_rng(9))
    base, _ = synth.simulate_reads([seq], k, mean, 150, seed=12, bases_per_read=430)
    n = 2300
    sigs = [base[i % len(base)] for i in range(n)]
    ids = [f"read_{i:06d}" for i in range(n)]
    fa, mf, b5 = str(tmp_path / "ref.fa"), str(tmp_path / "model.txt"), str(tmp_path / "reads.blow5")
    synth.write_fasta(fa, ["chrT"], [seq])
    synth.write_model_file(mf, k, mean, stdv)
    synth.write_blow5(b5, ids, sigs)
    for extra in ([], ["--sam"]):
        outs = []
        for K in ("4096", "200"):
            r = subprocess.run([cli, "dtw", fa, b5, "--kmer-model", mf, "--gpus", "1", "-t", "8", "-K", K, "-B", "100G"] + extra,
                               capture_output=True, text=True)
            assert r.returncode == 0, r.stderr[-2000:]
            outs.append(r.stdout)
        assert outs[0] == outs[1]
        body = [ln for ln in outs[0].splitlines() if not ln.startswith("@")]
        assert [ln.split("\t")[0] for ln in body] == ids


def test_cli_c4_scale_matches_reference_binary(cli, tmp_path):
    """BASELINE.json configs[3] shape through both command lines on the same files: one 1 Mb contig (R10 k=9,
    both strands, 5e8 cells per read).  The unmodified reference binary (oracle/_ref/sigfish, ~0.3 s per read
    on 16 cores) maps the first 16 reads; `sigfish-b200 dtw` maps 600 in several batches: the PAF of the common
    reads must be byte-identical."""
    if not H.have_ref_bin():
        pytest.skip("oracle/_ref not built")
    k = 9
    mean, stdv = synth.make_model(k)
    seq = synth.random_sequence(1_000_000, np.random.default_rng(1))
    sigs, _ = synth.simulate_reads([seq], k, mean, 600, seed=4242, bases_per_read=450)
    ids = [f"read_{i:06d}" for i in range(len(sigs))]
    fa, mf = str(tmp_path / "ref.fa"), str(tmp_path / "model.txt")
    synth.write_fasta(fa, ["chrS"], [seq])
    synth.write_model_file(mf, k, mean, stdv)
    synth.write_blow5(str(tmp_path / "all.blow5"), ids, sigs, kit="sqk-lsk114")
    synth.write_blow5(str(tmp_path / "head.blow5"), ids[:16], sigs[:16], kit="sqk-lsk114")
    r = subprocess.run([cli, "dtw", fa, str(tmp_path / "all.blow5"), "--kmer-model", mf, "--gpus", "1", "-K", "256"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "--pore r10 was set automatically" in r.stderr
    want = H.run_ref(fa, str(tmp_path / "head.blow5"), mf, threads=os.cpu_count() or 8)
    assert want.count("\n") == 16
    assert "".join(r.stdout.splitlines(keepends=True)[:16]) == want
    assert r.stdout.count("\n") == 600


@pytest.mark.parametrize("seed", range(10))
def test_cli_fuzz_matches_oracle_paf_and_sam(cli, tmp_path, seed):
    """seeded fuzz of the command line (C host + GPU) against the CPU oracle: random chemistry, flags, q, p
    (incl. -1), contigs and reads; PAF and SAM text must be identical"""
    rng = np.random.default_rng(8000 + seed)
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
    mean, stdv = synth.make_model(k, seed=41 + seed)
    seqs = [synth.random_sequence(int(n), rng) for n in rng.integers(700, 4000, size=int(rng.integers(1, 5)))]
    names = [f"c{i}" for i in range(len(seqs))]
    sigs, scs = [], []
    for r in range(7):
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
    c = dict(q=q, p=p, flags=flags)
    oflags = flags | (0x400 if rna004 else 0)
    out, err = _run(cli, c, fa, s5, mf, ["-K", "3"])
    if rna004:
        assert "--pore rna004 was set automatically" in err
    assert out == H.oracle_paf(names, seqs, mean, k, ids, sigs, scs, oflags, q, p), (flags, q, p)
    if not flags & H.F_DTW:
        out, _ = _run(cli, c, fa, s5, mf, ["--sam"])
        out = "".join(l for l in out.splitlines(keepends=True) if not l.startswith("@PG"))
        assert out == H.oracle_sam(names, seqs, mean, k, ids, sigs, scs, oflags, q, p), (flags, q, p)


# ---------------------------------------------------------------- the reference's own host code over libsfgpu.so

REF_ACC_BIN = os.path.join(H.ORACLE_DIR, "_ref", "sigfish_acc")
ACC_CASES = ["dna_sp1_default", "rna_sequin_default", "dna_synth48", "dna_sp1_from_end", "dna_short_reads", "dna_r10_k9",
             "rna_sequin_invert", "rna_sequin_dtw_std", "rna_sequin_q500_auto", "rna_tail24_auto", "rna004_tx2000_invert"]


def _run_acc(c, fa, reads, mf, extra=()):
    cmd = [REF_ACC_BIN, "dtw", fa, reads, "--kmer-model", mf, "-q", str(c["q"]), "-p", str(c["p"]), "-t", "4"] + \
        H.flags_to_cli(c["flags"]) + list(extra)
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout, r.stderr


@pytest.mark.parametrize("case", ACC_CASES)
def test_reference_binary_with_accelerator_seam_matches_golden(tmp_path, case):
    """integration/sigfish_acc.patch applied to the reference (oracle/Makefile: ref_acc) = the reference's `make acc=1`
    build with its dormant accelerator hooks bound to libsfgpu.so.  Its own main(), slow5lib loading, work_db(parse_single),
    paf_str() and output_db() run unchanged; events, query window, DTW and hit selection come from the GPU.  The PAF must
    equal what the same sources print on the CPU (tests/golden/paf)."""
    if not os.path.exists(REF_ACC_BIN):
        pytest.skip("oracle/_ref/sigfish_acc not built (needs /root/reference at build time)")
    c, fa, reads, mf = _inputs(str(tmp_path), case, "blow5")
    out, err = _run_acc(c, fa, reads, mf, ["--accel=yes"])
    assert "Initialising accelator" in err
    assert out == open(os.path.join(H.GOLDEN, "paf", case + ".paf")).read()


@pytest.mark.parametrize("case", ["dna_synth48", "rna_sequin_default", "rna_tail24_auto"])
def test_reference_binary_with_accelerator_seam_sam(tmp_path, case):
    """--sam through the patched reference: sfgpu_collect_paths() feeds the reference's own path_to_map() / sam_str()"""
    if not os.path.exists(REF_ACC_BIN):
        pytest.skip("oracle/_ref/sigfish_acc not built")
    c, fa, reads, mf = _inputs(str(tmp_path), case, "blow5")
    out, _ = _run_acc(c, fa, reads, mf, ["--accel=yes", "--sam"])
    want = open(os.path.join(H.GOLDEN, "sam", case + ".sam")).read()
    strip = lambda t: "".join(l for l in t.splitlines(keepends=True) if not l.startswith("@PG"))
    assert strip(out) == strip(want)


def test_reference_binary_accel_no_is_the_cpu_path(tmp_path):
    """the same patched binary with --accel=no runs the reference's CPU code: the seam is a switch, not a fork"""
    if not os.path.exists(REF_ACC_BIN):
        pytest.skip("oracle/_ref/sigfish_acc not built")
    c, fa, reads, mf = _inputs(str(tmp_path), "dna_sp1_default", "blow5")
    out, err = _run_acc(c, fa, reads, mf, ["--accel=no"])
    assert "Initialising accelator" not in err
    assert out == open(os.path.join(H.GOLDEN, "paf", "dna_sp1_default.paf")).read()
