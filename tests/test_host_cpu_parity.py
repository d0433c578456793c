"""The product's HOST code end to end on the CPU: sigfish_b200/host/*.c (readers, batch loop, read sharding over
contexts, device-decode fallback, threaded epilogue, PAF / SAM writers) linked against a TEST DOUBLE of libsfgpu.so
that answers include/sfgpu.h with the CPU oracle (tests/mockdev/sfgpu_oracle.c).  Its output must be, byte for byte,
what the unmodified reference binary printed (tests/golden/paf, tests/golden/sam) -- the same comparison
tests/test_gpu_cli.py makes with the CUDA library, minus the kernels.  Nothing here is part of the product: the double
is built into tests/mockdev/_build/ and only this file runs it."""
import glob
import gzip
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

import helpers as H
from sigfish_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
MOCK = os.path.join(HERE, "mockdev")
BUILD = os.path.join(MOCK, "_build")
CASES = json.load(open(os.path.join(H.GOLDEN, "cases.json")))
SAM_CASES = sorted(f[:-4] for f in os.listdir(os.path.join(H.GOLDEN, "sam")))
# cases the CPU oracle needs long for (2000 transcripts; 4.4 M reference columns per read with --full-ref): one of them
# runs here, the GPU suite runs them all
SLOW = {"rna004_tx2000_invert_full_ref", "rna004_tx2000_default", "rna004_tx2000_dtw_std"}
SLOW_SAM = SLOW | {"rna004_tx2000_invert", "rna004_tail16_auto"}


@pytest.fixture(scope="module")
def cli():
    H.build_oracle()
    os.makedirs(BUILD, exist_ok=True)
    inc = os.path.join(H.ROOT, "include")
    host = os.path.join(H.ROOT, "sigfish_b200", "host")
    lib = os.path.join(BUILD, "libsfgpu.so")
    exe = os.path.join(BUILD, "sigfish-b200-oracle-device")
    srcs = sorted(glob.glob(os.path.join(host, "*.c")))
    deps = srcs + glob.glob(os.path.join(host, "*.h")) + [os.path.join(MOCK, "sfgpu_oracle.c"), H.ORACLE_SO,
                                                           os.path.join(inc, "sfgpu.h")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.run(["gcc", "-O2", "-g", "-std=c99", "-Wall", "-fPIC", "-shared", "-I", inc, "-I", H.ORACLE_DIR, "-o", lib,
                        os.path.join(MOCK, "sfgpu_oracle.c"), "-L", H.ORACLE_DIR, "-loracle", "-Wl,-rpath," + H.ORACLE_DIR,
                        "-lz", "-lm"], check=True)
        subprocess.run(["gcc", "-O2", "-g", "-std=c99", "-Wall", "-D_GNU_SOURCE", "-I", inc, "-o", exe] + srcs +
                       ["-L", BUILD, "-lsfgpu", "-Wl,-rpath," + BUILD, "-Wl,-rpath," + H.ORACLE_DIR, "-lz", "-lpthread", "-lm"],
                       check=True)
    return exe


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
    reads = os.path.join(tmp, "reads." + fmt)
    if fmt == "slow5":
        synth.write_slow5_ascii(reads, ids, sigs, rna=rna, kit=H.case_kit(c), scalings=sc)
    else:
        synth.write_blow5(reads, ids, sigs, rna=rna, kit=H.case_kit(c), scalings=sc)
    mean, stdv = synth.make_model(c["k"])
    mf = os.path.join(tmp, "model.txt")
    synth.write_model_file(mf, c["k"], mean, stdv)
    return c, fa, reads, mf


def _run(cli, c, fa, reads, mf, extra=(), gpus=1, env=None):
    cmd = [cli, "dtw", fa, reads, "--kmer-model", mf, "-q", str(c["q"]), "-p", str(c["p"]), "--gpus", str(gpus)] + \
        H.flags_to_cli(c["flags"]) + list(extra)
    e = dict(os.environ, MOCK_GPUS=str(gpus))
    e.update(env or {})
    r = subprocess.run(cmd, capture_output=True, text=True, env=e)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout, r.stderr


def _strip_pg(t):
    return "".join(l for l in t.splitlines(keepends=True) if not l.startswith("@PG"))


@pytest.mark.parametrize("case", sorted(set(CASES) - SLOW))
def test_host_paf_is_byte_identical_to_reference(cli, tmp_path, case):
    fmt = "slow5" if sum(map(ord, case)) % 2 else "blow5"  # the other format than tests/test_gpu_cli.py uses for the case
    c, fa, reads, mf = _inputs(str(tmp_path), case, fmt)
    out, err = _run(cli, c, fa, reads, mf)
    assert out == open(os.path.join(H.GOLDEN, "paf", case + ".paf")).read()
    assert "total entries" in err


@pytest.mark.parametrize("case", [c for c in SAM_CASES if c not in SLOW_SAM])
def test_host_sam_is_byte_identical_to_reference(cli, tmp_path, case):
    c, fa, reads, mf = _inputs(str(tmp_path), case, "blow5")
    out, _ = _run(cli, c, fa, reads, mf, ["--sam", "-K", "5"])
    assert _strip_pg(out) == _strip_pg(open(os.path.join(H.GOLDEN, "sam", case + ".sam")).read())
    assert out.count("@PG\tID:sigfish") == 1


@pytest.mark.parametrize("case,extra,gpus,env", [
    ("dna_synth48", ["-K", "7", "-t", "3"], 1, {}),
    ("dna_synth48", ["-B", "30K"], 1, {}),
    ("dna_multi_contig", ["-K", "1"], 1, {}),
    ("dna_synth48", ["-K", "20"], 3, {}),                           # reads sharded over three contexts
    ("dna_synth48", [], 8, {}),
    ("rna_synth32", ["-K", "9", "-t", "4"], 2, {}),
    ("dna_synth48", ["--device-decode=no"], 2, {}),                 # records decoded by the host reader
    ("dna_synth48", ["-K", "11"], 2, {"MOCK_REJECT_RECORDS": "1"}),  # the device rejects the records: host fallback
    ("rna_tail24_auto", ["-K", "5"], 2, {"MOCK_REJECT_RECORDS": "1"}),
])
def test_host_batching_sharding_and_decode_paths_do_not_change_output(cli, tmp_path, case, extra, gpus, env):
    c, fa, reads, mf = _inputs(str(tmp_path), case, "blow5")
    out, err = _run(cli, c, fa, reads, mf, extra, gpus=gpus, env=env)
    assert out == open(os.path.join(H.GOLDEN, "paf", case + ".paf")).read()
    assert f"on {gpus} GPU(s)" in err
    if env:
        assert "decoding this batch on the host" in err


def test_host_sharded_sam_and_debug_break(cli, tmp_path):
    for case in ("rna_tail24_auto", "dna_synth48"):
        c, fa, reads, mf = _inputs(str(tmp_path), case, "blow5")
        out, _ = _run(cli, c, fa, reads, mf, ["--sam", "-K", "9"], gpus=2)
        assert _strip_pg(out) == _strip_pg(open(os.path.join(H.GOLDEN, "sam", case + ".sam")).read()), case
    c, fa, reads, mf = _inputs(str(tmp_path), "dna_synth48", "blow5")
    out, _ = _run(cli, c, fa, reads, mf, ["--debug-break", "1", "-K", "10"])  # stops after N + 1 batches (dtw_main.c:322-325)
    want = open(os.path.join(H.GOLDEN, "paf", "dna_synth48.paf")).read()
    assert out == "".join(want.splitlines(keepends=True)[:20])


def test_host_threaded_epilogue_on_large_batches(cli, tmp_path):
    """batches of >= 1024 reads run the epilogue on the -t threads in chunks of 512 reads: same bytes as small batches"""
    k = 6
    mean, stdv = synth.make_model(k)
    seq = synth.random_sequence(900, np.random.default_rng(9))
    base, _ = synth.simulate_reads([seq], k, mean, 40, seed=12, bases_per_read=220)
    n = 2300
    sigs = [base[i % len(base)] for i in range(n)]
    ids = [f"read_{i:06d}" for i in range(n)]
    fa, mf, b5 = str(tmp_path / "ref.fa"), str(tmp_path / "model.txt"), str(tmp_path / "reads.blow5")
    synth.write_fasta(fa, ["chrT"], [seq])
    synth.write_model_file(mf, k, mean, stdv)
    synth.write_blow5(b5, ids, sigs)
    c = dict(q=50, p=20, flags=0)
    for extra in ([], ["--sam"]):
        a, _ = _run(cli, c, fa, b5, mf, ["-t", "8", "-K", "4096", "-B", "100G"] + extra, gpus=2)
        b, _ = _run(cli, c, fa, b5, mf, ["-t", "2", "-K", "200", "-B", "100G"] + extra)
        assert a == b
        assert [ln.split("\t")[0] for ln in a.splitlines() if not ln.startswith("@")] == ids


def _stable_stderr(text):
    import re
    ansi = re.compile(r"\x1b\[[0-9;]*m")
    keep = ("total entries", "total bytes", "Detected", "Autodetect query start")
    return [ansi.sub("", l).strip() for l in text.splitlines() if any(k in l for k in keep)]


def _fuzz_seeds():
    """12 seeds in the suite; SF_FUZZ_SEEDS=first:last runs another range (long runs are recorded in DESIGN.md 3)"""
    r = os.environ.get("SF_FUZZ_SEEDS", "")
    if ":" in r:
        a, b = r.split(":")
        return range(int(a), int(b))
    return range(12)


@pytest.mark.refbin
@pytest.mark.parametrize("seed", _fuzz_seeds())
def test_host_fuzz_matches_reference_binary(cli, tmp_path, seed):
    """seeded fuzz of the command line over the oracle device against the unmodified reference binary run right now
    (build container only): chemistry, flags, q, p (incl. -1), contigs, reads, batch size, contexts; PAF and SAM"""
    if not H.have_ref_bin():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(9000 + seed)
    rna = bool(rng.integers(0, 2))
    k = (9 if seed % 3 == 2 else 5) if rna else int(rng.choice([6, 9]))
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
    mean, stdv = synth.make_model(k, seed=51 + seed)
    seqs = [synth.random_sequence(int(n), rng) for n in rng.integers(700, 4000, size=int(rng.integers(1, 5)))]
    names = [f"c{i}" for i in range(len(seqs))]
    sigs, scs = [], []
    for r in range(9):
        if rna and p < 0 and r % 2 == 0:
            s, _ = synth.simulate_rna_reads_with_tail(seqs, k, mean, 1, seed=int(rng.integers(1 << 30)), bases_per_read=500)
        else:
            s, _ = synth.simulate_reads(seqs, k, mean, 1, seed=int(rng.integers(1 << 30)), rna=rna,
                                        bases_per_read=int(rng.choice([150, 300, 450, 700])), min_samples=700)
        sigs.append(s[0])
        scs.append(synth.RNA_SCALING if rna else synth.DNA_SCALING)
    ids = [f"r{i}" for i in range(len(sigs))]
    fa, mf = str(tmp_path / "ref.fa"), str(tmp_path / "model.txt")
    fmt = "slow5" if seed % 4 == 3 else "blow5"
    s5 = str(tmp_path / ("reads." + fmt))
    synth.write_fasta(fa, names, seqs)
    kit = "sqk-rna004" if rna004 else ("sqk-lsk114" if k == 9 else None)
    if fmt == "slow5":
        synth.write_slow5_ascii(s5, ids, sigs, rna=rna, kit=kit, scalings=scs)
    else:
        synth.write_blow5(s5, ids, sigs, rna=rna, kit=kit, scalings=scs, record_zlib=bool(seed % 2), signal_svb=bool(seed % 3))
    synth.write_model_file(mf, k, mean, stdv)
    c = dict(q=q, p=p, flags=flags)
    gpus = int(rng.integers(1, 4))
    extra = ["-K", str(int(rng.integers(1, 6))), "-t", str(int(rng.integers(1, 5)))]
    out, err = _run(cli, c, fa, s5, mf, extra, gpus=gpus)
    assert out == H.run_ref(fa, s5, mf, flags=flags, q=q, p=p), (flags, q, p, gpus, extra)
    # the lines of stderr that do not hold times: chemistry detection, the summary counters
    r = subprocess.run([H.REF_BIN, "dtw", fa, s5, "--kmer-model", mf, "-q", str(q), "-p", str(p)] + H.flags_to_cli(flags),
                       capture_output=True, text=True)
    assert _stable_stderr(err) == _stable_stderr(r.stderr), (flags, q, p)
    if not flags & H.F_DTW:  # the reference aborts on --dtw-std --sam (sigfish.c:669)
        out, _ = _run(cli, c, fa, s5, mf, extra + ["--sam"], gpus=gpus)
        want = H.run_ref(fa, s5, mf, flags=flags, q=q, p=p, extra=["--sam"])
        assert _strip_pg(out) == _strip_pg(want), (flags, q, p, gpus, extra)


# ---------------------------------------------------------------- the reference's own host code over the test double

REF_ACC_BIN = os.path.join(H.ORACLE_DIR, "_ref", "sigfish_acc")
ACC_CASES = ["dna_sp1_default", "rna_sequin_default", "dna_synth48", "dna_sp1_from_end", "dna_short_reads", "dna_r10_k9",
             "rna_sequin_invert", "rna_sequin_dtw_std", "rna_sequin_q500_auto", "rna_tail24_auto"]


@pytest.mark.refbin
@pytest.mark.parametrize("case", ACC_CASES)
def test_patched_reference_over_the_test_double_matches_golden(cli, tmp_path, case):
    """integration/sigfish_acc.patch (the reference built with its accelerator hooks bound to sfgpu.h,
    oracle/_ref/sigfish_acc) with the test double in place of libsfgpu.so: the glue of the patch (acc_b200.h: submit,
    collect, the fields it writes back into db, paf_str / sam_str on them) reproduces the golden PAF / SAM on the CPU.
    The GPU suite runs the same binary over the CUDA library."""
    if not os.path.exists(REF_ACC_BIN):
        pytest.skip("oracle/_ref/sigfish_acc not built (needs /root/reference at build time)")
    c, fa, reads, mf = _inputs(str(tmp_path), case, "blow5")
    env = dict(os.environ, LD_LIBRARY_PATH=BUILD)  # the binary finds libsfgpu.so through a RUNPATH: this comes first
    for sam in (False, True):
        if sam and (case not in SAM_CASES or c["flags"] & H.F_DTW):
            continue
        cmd = [REF_ACC_BIN, "dtw", fa, reads, "--kmer-model", mf, "-q", str(c["q"]), "-p", str(c["p"]), "-t", "4"] + \
            H.flags_to_cli(c["flags"]) + (["--sam"] if sam else [])
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr[-2000:]
        want = open(os.path.join(H.GOLDEN, "sam" if sam else "paf", case + (".sam" if sam else ".paf"))).read()
        assert _strip_pg(r.stdout) == _strip_pg(want), (case, sam)  # (without the double there is no device: exit 1)


def _fasta_variants(names, seqs, rng):
    """the same contigs written the ways kseq accepts them (reference src/kseq.h:184-224)"""
    def wrap(s, w):
        return [s[i:i + w] for i in range(0, len(s), w)]
    v = {}
    v["crlf_comments_blank_lines"] = b"".join(
        b">" + n.encode() + b" some comment\tmore\r\n" + b"\r\n".join(wrap(s, 61)) + b"\r\n\r\n" for n, s in zip(names, seqs))
    v["lowercase_and_n"] = b"".join(b">" + n.encode() + b"\n" + b"\n".join(wrap(s[:200].lower() + b"NNNNRY" + s[206:], 70)) + b"\n"
                                    for n, s in zip(names, seqs))
    v["text_before_first_record_no_final_newline"] = b"junk line\nmore junk\n" + b"".join(
        b">" + n.encode() + b"\n" + s + b"\n" for n, s in zip(names, seqs))[:-1]
    fq = b""
    for n, s in zip(names, seqs):
        q = bytes(rng.integers(33, 74, len(s), dtype=np.uint8))
        q = b"@" + q[1:60] + b">" + q[61:]  # quality lines that start with header characters: skipped by length
        fq += b"@" + n.encode() + b" desc\n" + b"\n".join(wrap(s, 60)) + b"\n+" + n.encode() + b"\n" + b"\n".join(wrap(q, 60)) + b"\n"
    v["fastq_multiline"] = fq
    v["fastq_then_truncated_quality"] = fq + b"@late\nACGTACGTACGTACGT\n+\nIIII\n"  # ends the file, no record
    v["mixed_fasta_fastq"] = fq[:fq.index(b"@" + names[1].encode())] + b"".join(
        b">" + n.encode() + b"\n" + s + b"\n" for n, s in zip(names[1:], seqs[1:])) if len(names) > 1 else fq
    return v


@pytest.mark.refbin
def test_host_reads_fasta_and_fastq_like_the_reference(cli, tmp_path):
    """the reference file as kseq would read it: Windows line ends, comments, blank lines, lower case and non-ACGT
    bases, text before the first record, FASTQ (multi-line, quality lines starting with '@' / '>'), a truncated
    FASTQ record at the end, gzip -- the PAF (contig names, lengths, coordinates) must equal the reference binary's"""
    if not H.have_ref_bin():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(77)
    k = 6
    mean, stdv = synth.make_model(k)
    seqs = [synth.random_sequence(int(n), rng) for n in (2500, 1800, 900)]
    names = ["chrA", "chrB.1", "c|3"]
    sigs, _ = synth.simulate_reads(seqs, k, mean, 8, seed=5, bases_per_read=400)
    ids = [f"r{i}" for i in range(len(sigs))]
    mf, s5 = str(tmp_path / "model.txt"), str(tmp_path / "reads.blow5")
    synth.write_model_file(mf, k, mean, stdv)
    synth.write_blow5(s5, ids, sigs)
    c = dict(q=250, p=50, flags=0)
    outs = {}
    for name, data in _fasta_variants(names, seqs, rng).items():
        for gz in (False, True):
            fa = str(tmp_path / (name + (".fa.gz" if gz else ".fa")))
            with (gzip.open(fa, "wb") if gz else open(fa, "wb")) as f:
                f.write(data)
            out, _ = _run(cli, c, fa, s5, mf)
            assert out == H.run_ref(fa, s5, mf), (name, gz)
            outs[name] = out
    assert outs["fastq_multiline"] == outs["fastq_then_truncated_quality"] == outs["crlf_comments_blank_lines"]
    assert outs["fastq_multiline"].count("\n") == len(ids)


@pytest.mark.refbin
def test_host_option_errors_are_the_reference_binarys(cli, tmp_path):
    """option checking of `dtw` (reference src/dtw_main.c:140-280): same exit status and the same ERROR lines (the
    reference appends the source position) as the reference binary run right now"""
    if not H.have_ref_bin():
        pytest.skip("oracle/_ref not built")
    import re
    c, fa, reads, mf = _inputs(str(tmp_path), "dna_synth48", "blow5")
    ansi = re.compile(r"\x1b\[[0-9;]*m")

    def errors(cmd, env=None):
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        lines = [ansi.sub("", l) for l in r.stderr.splitlines()]
        keep = [re.sub(r"\s+At \S+:\d+$", "", l).strip() for l in lines if "::ERROR]" in l or "unrecognized option" in l or
                "requires an argument" in l or "invalid option" in l]
        return r.returncode != 0, keep

    for extra in (["--dtw-std"], ["--invert"], ["--full-ref"], ["-p", "-1"], ["--pore", "r11"], ["-q", "-5"], ["-K", "0"],
                  ["-t", "0"], ["-B", "0"], ["--rna", "-p", "-1", "--from-end"], ["--rna", "-p", "-1", "--invert"],
                  ["--bogus"], ["-K"], ["--sam", "--dtw-std"]):
        want = errors([H.REF_BIN, "dtw", fa, reads, "--kmer-model", mf] + extra)
        got = errors([cli, "dtw", fa, reads, "--kmer-model", mf, "--gpus", "1"] + extra, dict(os.environ, MOCK_GPUS="1"))
        assert got == want, (extra, got, want)  # (an unrecognised option is reported and ignored by both)


@pytest.mark.refbin
@pytest.mark.parametrize("reads,fasta,k,args", [
    ("sp1_dna", "nCoV-2019", 6, []),
    ("sp1_dna", "nCoV-2019", 6, ["--from-end"]),
    ("sequin_rna", "rnasequin", 5, ["--rna", "-q", "500", "-p", "-1"]),
    ("sequin_rna", "rnasequin", 5, ["--rna", "--full-ref", "-q", "500", "-p", "-1"]),
    ("sequin_rna", "rnasequin", 5, ["--rna", "--from-end", "-q", "500"]),
    ("sequin_rna", "rnasequin", 5, ["--rna", "--full-ref", "--from-end", "-q", "500"]),
])
def test_host_runs_the_reference_test_script_commands(cli, tmp_path, reads, fasta, k, args):
    """the six `sigfish dtw` commands of the reference's test/test.sh and test/test_extensive.sh (lines 58-86) on its
    own reads and references (synthetic pore model: the built-in tables are absent from the mount), both binaries
    run right now with the script's -t 8: the PAF must be identical"""
    if not H.have_ref_bin():
        pytest.skip("oracle/_ref not built")
    ids, sigs, sc = H.load_reads_npz(os.path.join(H.GOLDEN, reads + ".npz"))
    s5, fa, mf = str(tmp_path / "reads.blow5"), str(tmp_path / "ref.fa"), str(tmp_path / "model.txt")
    synth.write_blow5(s5, ids, sigs, rna="--rna" in args, scalings=sc)
    with gzip.open(os.path.join(H.GOLDEN, fasta + ".fa.gz"), "rb") as fi, open(fa, "wb") as fo:
        shutil.copyfileobj(fi, fo)
    mean, stdv = synth.make_model(k)
    synth.write_model_file(mf, k, mean, stdv)
    want = subprocess.run([H.REF_BIN, "dtw", fa, s5, "--kmer-model", mf, "-t", "8"] + args, capture_output=True, text=True)
    assert want.returncode == 0, want.stderr[-2000:]
    got = subprocess.run([cli, "dtw", fa, s5, "--kmer-model", mf, "-t", "8", "--gpus", "2"] + args, capture_output=True, text=True,
                         env=dict(os.environ, MOCK_GPUS="2"))
    assert got.returncode == 0, got.stderr[-2000:]
    assert got.stdout == want.stdout and got.stdout.count("\n") == len(ids)
    assert _stable_stderr(got.stderr) == _stable_stderr(want.stderr)


@pytest.mark.refbin
@pytest.mark.parametrize("fasta,blow5,k,args", [
    ("nCoV-2019.reference.fasta", "sp1_dna.blow5", 6, []),
    ("rnasequin_sequences_2.4.fa", "sequin_rna.blow5", 5, ["--rna", "-q", "500", "-p", "-1"]),
])
def test_host_on_the_reference_files_as_they_are(cli, tmp_path, fasta, blow5, k, args):
    """test/test.sh's two commands on the reference's files themselves (slow5tools-written BLOW5 with auxiliary fields,
    its FASTA files), records path and host-decode path"""
    d = "/root/reference/test"
    if not (H.have_ref_bin() and os.path.exists(os.path.join(d, blow5))):
        pytest.skip("reference mount / oracle/_ref not present")
    mf = str(tmp_path / "model.txt")
    mean, stdv = synth.make_model(k)
    synth.write_model_file(mf, k, mean, stdv)
    base = ["dtw", os.path.join(d, fasta), os.path.join(d, blow5), "--kmer-model", mf, "-t", "8"] + args
    want = subprocess.run([H.REF_BIN] + base, capture_output=True, text=True)
    assert want.returncode == 0 and want.stdout.count("\n") >= 5
    for extra in ([], ["--device-decode=no"], ["--sam"]):
        got = subprocess.run([cli] + base + ["--gpus", "2"] + extra, capture_output=True, text=True, env=dict(os.environ, MOCK_GPUS="2"))
        assert got.returncode == 0, got.stderr[-2000:]
        if "--sam" in extra:
            ref_sam = subprocess.run([H.REF_BIN] + base + ["--sam"], capture_output=True, text=True)
            assert _strip_pg(got.stdout) == _strip_pg(ref_sam.stdout)
        else:
            assert got.stdout == want.stdout


@pytest.mark.refbin
def test_host_misc_options_behave_like_the_reference(cli, tmp_path):
    """options that do not change the mapping (or only where it goes): same exit status, same stdout, same -o file"""
    if not H.have_ref_bin():
        pytest.skip("oracle/_ref not built")
    c, fa, reads, mf = _inputs(str(tmp_path), "dna_sp1_default", "blow5")
    out_r, out_h = str(tmp_path / "r.paf"), str(tmp_path / "h.paf")
    for extra in (["--accel=yes"], ["--accel=no"], ["--profile-cpu=yes"], ["--profile-cpu=maybe"], ["--verbose", "0"],
                  ["--verbose", "6"], ["-o", "FILE"], ["-o", "-"], ["--secondary=yes"], ["-a"], ["-w", "chr1:1-100"],
                  ["--kmer-model", "nofile.txt"], ["-K", "2", "--debug-break=yes"]):
        ex_r = [out_r if x == "FILE" else x for x in extra]
        ex_h = [out_h if x == "FILE" else x for x in extra]
        want = subprocess.run([H.REF_BIN, "dtw", fa, reads, "--kmer-model", mf] + ex_r, capture_output=True, text=True)
        got = subprocess.run([cli, "dtw", fa, reads, "--kmer-model", mf, "--gpus", "1"] + ex_h, capture_output=True, text=True,
                             env=dict(os.environ, MOCK_GPUS="1"))
        assert got.returncode == want.returncode, extra
        assert _strip_pg(got.stdout) == _strip_pg(want.stdout), extra
        if "FILE" in extra:
            assert open(out_h).read() == open(out_r).read() and got.stdout == ""


def test_host_refuses_contigs_shorter_than_k(cli, tmp_path):
    """a contig with fewer bases than the k-mer size: the reference binary dies with a segmentation fault (gen_ref
    computes a negative length); the library refuses it (SFGPU_EARG) and the command line says so and exits 1"""
    c, fa, reads, mf = _inputs(str(tmp_path), "dna_sp1_default", "blow5")
    for head in (b">empty\n", b">tiny\nACG\n"):
        bad = str(tmp_path / "bad.fa")
        open(bad, "wb").write(head + open(fa, "rb").read())
        r = subprocess.run([cli, "dtw", bad, reads, "--kmer-model", mf, "--gpus", "1"], capture_output=True, text=True,
                           env=dict(os.environ, MOCK_GPUS="1"))
        assert r.returncode == 1 and "ERROR" in r.stderr and r.stdout == ""


@pytest.mark.parametrize("fmt,extra", [("blow5", []), ("blow5", ["--device-decode=no"]), ("slow5", [])])
def test_host_empty_reads_print_nothing(cli, tmp_path, fmt, extra):
    """reads without samples (reference src/sigfish.c:1068-1070: nothing is printed for them; its binary crashes on
    some such files) between ordinary reads: the other rows are what they are without the empty reads"""
    c, fa, reads, mf = _inputs(str(tmp_path), "dna_sp1_default", "blow5")
    ids, sigs, sc = H.case_reads(CASES["dna_sp1_default"])
    ids2 = ["e0"] + list(ids[:2]) + ["e1"] + list(ids[2:]) + ["e2"]
    z = np.zeros(0, dtype=np.int16)
    sigs2 = [z] + list(sigs[:2]) + [z] + list(sigs[2:]) + [z]
    sc2 = [sc[0]] + list(sc[:2]) + [sc[0]] + list(sc[2:]) + [sc[0]]
    p = str(tmp_path / ("with_empty." + fmt))
    (synth.write_blow5 if fmt == "blow5" else synth.write_slow5_ascii)(p, ids2, sigs2, scalings=sc2)
    out, err = _run(cli, c, fa, p, mf, extra + ["-K", "3"], gpus=2)
    assert out == open(os.path.join(H.GOLDEN, "paf", "dna_sp1_default.paf")).read()
    assert f"total entries: {len(ids2)}" in err
