"""CPU-side behaviour of the `sigfish-b200` command line: everything that happens before a GPU is needed
(usage, version, option validation in the reference's order) and the loud failure when there is no device."""
import os
import subprocess

import numpy as np
import pytest

import helpers as H
from sigfish_b200 import build as B
from sigfish_b200 import capi, synth


@pytest.fixture(scope="module")
def cli():
    B.build_all()
    return B.CLI


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp("cli")
    k = 5
    mean, stdv = synth.make_model(k)
    rng = np.random.default_rng(0)
    seqs = [synth.random_sequence(900, rng)]
    sigs, _ = synth.simulate_reads(seqs, k, mean, 2, seed=1, rna=True, bases_per_read=300)
    fa, s5, mf = str(d / "r.fa"), str(d / "r.blow5"), str(d / "m.txt")
    synth.write_fasta(fa, ["t1"], seqs)
    synth.write_blow5(s5, ["a", "b"], sigs, rna=True)
    synth.write_model_file(mf, k, mean, stdv)
    return fa, s5, mf


def run(cli, *args):
    return subprocess.run([cli, *args], capture_output=True, text=True)


def test_usage_version_help(cli):
    r = run(cli)
    assert r.returncode != 0 and "Usage: sigfish-b200 <command>" in r.stderr
    r = run(cli, "--version")
    assert r.returncode == 0 and r.stdout.startswith("sigfish ")
    r = run(cli, "dtw", "--version")
    assert r.returncode == 0 and r.stdout.startswith("sigfish ")
    r = run(cli, "dtw", "-h")
    assert r.returncode == 0 and "Usage: sigfish dtw [OPTIONS] genome.fa reads.blow5" in r.stdout
    for opt in ("-K INT", "-B FLOAT", "--kmer-model", "--rna", "-q INT", "-p INT", "--dtw-std", "--invert", "--full-ref",
                "--from-end", "--sam", "--pore", "--gpus"):
        assert opt in r.stdout, opt
    r = run(cli, "real")  # a sub-tool of other sigfish versions that is outside this path
    assert r.returncode != 0 and "Unrecognised command" in r.stderr


def test_option_validation_precedes_everything(cli, files):
    fa, s5, mf = files
    cases = [(["-K", "0"], "Batch size should larger than 0"), (["-t", "0"], "Number of threads should larger than 0"),
             (["-B", "0"], "Maximum number of bytes should be larger than 0"), (["-q", "-5"], "Query size should larger than 0"),
             (["--gpus", "0"], "Number of GPUs should larger than 0"),
             (["--rna", "-p", "-1", "--invert"], "Inversion is not compatible with auto query start"),
             (["--rna", "-p", "-1", "--from-end"], "not compatible with auto query start"),
             (["--dtw-std"], "DTW is only available for RNA")]  # the file IS RNA, but the check precedes detection
    for extra, msg in cases:
        r = run(cli, "dtw", fa, s5, "--kmer-model", mf, *extra)
        assert r.returncode != 0 and msg in r.stderr, (extra, r.stderr[-400:])
    r = run(cli, "dtw", fa, s5, "--kmer-model", mf, "--pore", "rna004", "-K", "0")  # accepted here (reference bug F6)
    assert "Pore model should be" not in r.stderr


def test_no_gpu_is_a_hard_error(cli, files):
    if capi.lib().sfgpu_device_count() > 0:
        pytest.skip("a GPU is present")
    fa, s5, mf = files
    r = run(cli, "dtw", fa, s5, "--kmer-model", mf)
    assert r.returncode != 0
    assert "no sm_100 (B200) GPU visible" in r.stderr and "no CPU path" in r.stderr
    assert r.stdout == ""
    # a missing model is reported before the device is looked for
    r = run(cli, "dtw", fa, s5)
    assert r.returncode != 0 and "--kmer-model" in r.stderr


EVAL = os.path.join(H.GOLDEN, "eval")
EVAL_CASES = {
    "dna_synth48_vs_from_end": ("dna_synth48.paf", "dna_synth48_from_end.paf", []),
    "rna_default_vs_full_ref": ("rna_sequin_default.paf", "rna_sequin_full_ref.paf", []),
    "rna_default_vs_full_ref_tid_only": ("rna_sequin_default.paf", "rna_sequin_full_ref.paf", ["--tid-only"]),
    "disjoint_read_sets": ("dna_short_reads.paf", "dna_synth48.paf", []),
    "crafted": ("crafted_truth.paf", "crafted_test.paf", []),
    "crafted_no_secondary": ("crafted_truth.paf", "crafted_test.paf", ["--secondary", "no"]),
    "crafted_tid_only": ("crafted_truth.paf", "crafted_test.paf", ["--tid-only"]),
}


@pytest.mark.parametrize("case", sorted(EVAL_CASES))
def test_eval_report_matches_reference_golden(cli, case):
    """`sigfish-b200 eval` prints the report the reference's `sigfish eval` printed for the same PAF pair
    (tests/golden/eval/*.txt, made by tests/golden/make_golden.py)"""
    truth, test, opts = EVAL_CASES[case]
    loc = lambda f: os.path.join(EVAL, f) if f.startswith("crafted") else os.path.join(H.GOLDEN, "paf", f)
    r = run(cli, "eval", *opts, loc(truth), loc(test))
    assert r.returncode == 0, r.stderr
    assert r.stdout == open(os.path.join(EVAL, case + ".txt")).read()
    assert "Total mappings in testset" in r.stderr


def test_eval_usage_and_errors(cli, tmp_path):
    r = run(cli, "eval")
    assert r.returncode != 0 and "Usage: sigfish eval truth.paf test.paf" in r.stderr
    r = run(cli, "eval", "-h")
    assert r.returncode == 0 and "--tid-only" in r.stdout
    r = run(cli, "eval", str(tmp_path / "nope.paf"), str(tmp_path / "nope2.paf"))
    assert r.returncode != 0 and "cannot open" in r.stderr
    bad = tmp_path / "bad.paf"
    bad.write_text("r1\t100\t0\t50\t+\tchr\n")
    r = run(cli, "eval", str(bad), str(bad))
    assert r.returncode != 0 and "malformed PAF" in r.stderr


@pytest.mark.refbin
def test_patched_reference_keeps_its_cpu_path_and_fails_loudly_without_a_gpu(tmp_path):
    """integration/sigfish_acc.patch, applied and compiled by oracle/Makefile (ref_acc): with --accel=no the binary
    is the reference (golden PAF from the CPU code); with --accel=yes and no device it exits with the library's
    error instead of falling back."""
    import json
    acc = os.path.join(H.ORACLE_DIR, "_ref", "sigfish_acc")
    if not os.path.exists(acc):
        pytest.skip("oracle/_ref/sigfish_acc not built (needs /root/reference at build time)")
    patch = open(os.path.join(H.ROOT, "integration", "sigfish_acc.patch")).read()
    for sym in ("sfgpu_create", "sfgpu_set_ref", "sfgpu_submit_reads", "sfgpu_collect", "sfgpu_collect_paths", "sfgpu_destroy"):
        assert sym in patch
    c = json.load(open(os.path.join(H.GOLDEN, "cases.json")))["dna_sp1_default"]
    names, seqs = H.case_fasta(c)
    ids, sigs, sc = H.case_reads(c)
    fa, s5, mf = str(tmp_path / "ref.fa"), str(tmp_path / "reads.blow5"), str(tmp_path / "model.txt")
    synth.write_fasta(fa, names, seqs)
    synth.write_blow5(s5, ids, sigs, scalings=sc)
    synth.write_model_file(mf, c["k"], *synth.make_model(c["k"]))
    r = subprocess.run([acc, "dtw", fa, s5, "--kmer-model", mf, "--accel=no"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-1000:]
    assert r.stdout == open(os.path.join(H.GOLDEN, "paf", "dna_sp1_default.paf")).read()
    if capi.lib().sfgpu_device_count() == 0:
        r = subprocess.run([acc, "dtw", fa, s5, "--kmer-model", mf, "--accel=yes"], capture_output=True, text=True)
        assert r.returncode != 0 and "no CPU fallback" in r.stderr and r.stdout == ""


@pytest.mark.refbin
def test_eval_fuzz_matches_reference_binary(cli, tmp_path):
    """random truth / test PAF pairs (secondary rows, shifted and missing mappings, shuffled order) through
    `sigfish-b200 eval` and the reference's `sigfish eval` run right now, every option: identical reports"""
    if not H.have_ref_bin():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(3)

    def rows(reads, contigs, sec_p, miss_p):
        out = []
        for r in reads:
            if rng.random() < miss_p:
                continue
            for j in range(1 + int(rng.random() < sec_p) * int(rng.integers(1, 4))):
                st = int(rng.integers(0, 20000))
                en = st + int(rng.integers(1, 400))
                out.append(f"{r}\t{int(rng.integers(1000, 9000))}\t{int(rng.integers(0, 300))}\t{int(rng.integers(300, 2000))}\t"
                           f"{'+-'[int(rng.integers(0, 2))]}\t{contigs[int(rng.integers(0, len(contigs)))]}\t30000\t{st}\t{en}\t"
                           f"{int(rng.integers(50, 200))}\t{en - st}\t{int(rng.integers(0, 61))}\ttp:A:{'P' if j == 0 else 'S'}")
        return out

    t_path, q_path = str(tmp_path / "truth.paf"), str(tmp_path / "test.paf")
    for it in range(48):
        reads = [f"read_{i}" for i in range(int(rng.integers(1, 40)))]
        contigs = [f"chr{i}" for i in range(int(rng.integers(1, 4)))]
        truth = rows(reads, contigs, 0.3, 0.1)
        test = []
        for row in truth:
            u, f = rng.random(), row.split("\t")
            if u < 0.5:
                test.append(row)
            elif u < 0.8:
                d = int(rng.integers(-300, 300))
                f[7] = str(max(0, int(f[7]) + d))
                f[8] = str(max(int(f[7]) + 1, int(f[8]) + d))
                test.append("\t".join(f))
        test += rows(reads, contigs, 0.2, 0.5)
        if rng.random() < 0.3:
            test = [test[i] for i in rng.permutation(len(test))]
        open(t_path, "w").write("".join(l + "\n" for l in truth))
        open(q_path, "w").write("".join(l + "\n" for l in test))
        opts = [[], ["--tid-only"], ["--secondary", "no"], ["--secondary", "yes"]][it % 4]
        want = subprocess.run([H.REF_BIN, "eval"] + opts + [t_path, q_path], capture_output=True, text=True)
        got = run(cli, "eval", *opts, t_path, q_path)
        assert (got.returncode != 0) == (want.returncode != 0), (it, opts)
        assert got.stdout == want.stdout, (it, opts)


@pytest.mark.refbin
@pytest.mark.parametrize("truth,test", [("sp1_dna.minimap2.paf", "dna_sp1_default.paf"),
                                        ("sequin_rna.minimap2.paf", "rna_sequin_q500_auto.paf")])
def test_eval_on_the_reference_truth_sets(cli, truth, test):
    """the EVALUATE step of the reference's test/test.sh:23-42: its minimap2 truth sets against the mappings of its two
    test commands (golden PAF; made with a synthetic pore model, so `correct` is low -- the comparison is between the
    two eval implementations, not against the script's thresholds, which need the built-in pore tables)"""
    t = os.path.join("/root/reference/test", truth)
    if not (H.have_ref_bin() and os.path.exists(t)):
        pytest.skip("reference mount / oracle/_ref not present")
    q = os.path.join(H.GOLDEN, "paf", test)
    for opts in ([], ["--tid-only"], ["--secondary", "no"]):
        want = subprocess.run([H.REF_BIN, "eval"] + opts + [t, q], capture_output=True, text=True)
        got = run(cli, "eval", *opts, t, q)
        assert want.returncode == 0 and got.returncode == 0
        assert got.stdout == want.stdout and "mapped_testset" in got.stdout
