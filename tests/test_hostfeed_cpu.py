"""CPU tests of the CLI's threaded host pipeline (loader -> decoder -> submit / collect / output,
sigfish_b200/host/dtw_cli.c) against the NULL device of tools/hostfeed: no alignment happens, the stub
returns made-up hits, so only the host logic is checked -- every record is decoded, batches of any size
come out complete and in input order, the summary counters add up."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools", "hostfeed"))
import run as hostfeed  # noqa: E402
from sigfish_b200 import synth  # noqa: E402


@pytest.fixture(scope="module")
def feed(tmp_path_factory):
    exe = hostfeed.build()
    d = tmp_path_factory.mktemp("hostfeed")
    k = 6
    mean, stdv = synth.make_model(k)
    seq = synth.random_sequence(5000, np.random.default_rng(3))
    sigs, _ = synth.simulate_reads([seq], k, mean, 61, seed=8, bases_per_read=400)
    ids = [f"read_{i:04d}" for i in range(len(sigs))]
    synth.write_model_file(str(d / "model.txt"), k, mean, stdv)
    synth.write_fasta(str(d / "ref.fa"), ["chrS"], [seq])
    synth.write_blow5(str(d / "reads.blow5"), ids, sigs)
    synth.write_blow5(str(d / "empty.blow5"), [], [])
    return exe, d, ids, sigs


def run_cli(exe, d, blow5, *extra, gpus=2):
    env = dict(os.environ, HOSTFEED_GPUS=str(gpus))
    r = subprocess.run([exe, "dtw", str(d / "ref.fa"), str(d / blow5), "--kmer-model", str(d / "model.txt"), "-t", "3",
                        "--gpus", str(gpus)] + list(extra), capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.splitlines(), r.stderr


@pytest.mark.parametrize("K", [1, 2, 7, 60, 61, 62, 1000])
def test_batches_come_out_complete_and_in_order(feed, K):
    exe, d, ids, sigs = feed
    lines, err = run_cli(exe, d, "reads.blow5", "-K", str(K), "-B", "100G")
    assert [ln.split("\t")[0] for ln in lines] == ids
    # column 2 of a PAF line is the raw signal length: every record went through the decoder
    assert [int(ln.split("\t")[1]) for ln in lines] == [len(s) for s in sigs]
    m = re.search(r"total entries: (\d+)", err)
    assert m and int(m.group(1)) == len(ids)
    n_batches = len(re.findall(r"Entries \(.*\) loaded", err))
    assert n_batches == len(ids) // K + 1 if len(ids) % K == 0 else n_batches == -(-len(ids) // K)


def test_byte_limit_ends_batches_too(feed):
    exe, d, ids, _ = feed
    lines, err = run_cli(exe, d, "reads.blow5", "-K", "1000", "-B", "20K")
    assert [ln.split("\t")[0] for ln in lines] == ids
    assert len(re.findall(r"Entries \(.*\) loaded", err)) > 3


def test_empty_file_and_one_gpu(feed):
    exe, d, ids, _ = feed
    lines, err = run_cli(exe, d, "empty.blow5")
    assert lines == [] and "total entries: 0" in err
    lines, _ = run_cli(exe, d, "reads.blow5", "-K", "5", gpus=1)
    assert [ln.split("\t")[0] for ln in lines] == ids


def test_debug_break_stops_after_the_given_batch(feed):
    exe, d, ids, _ = feed
    lines, _ = run_cli(exe, d, "reads.blow5", "-K", "10", "--debug-break", "1")
    assert [ln.split("\t")[0] for ln in lines] == ids[:20]  # batches 0 and 1 (src/dtw_main.c:322-325)


def test_threaded_epilogue_keeps_order_and_counters(feed, tmp_path):
    """batches of >= 1024 reads: the per-read epilogue (PAF text, counters) runs on the -t worker threads in chunks of
    512 reads; lines must still come out in input order, one per read, and the summary must add up"""
    exe, d, ids, sigs = feed
    n = 2600
    big_ids = [f"big_{i:05d}" for i in range(n)]
    big_sigs = [sigs[i % len(sigs)] for i in range(n)]
    synth.write_blow5(str(d / "big.blow5"), big_ids, big_sigs)
    for K in ("2600", "1100"):
        lines, err = run_cli(exe, d, "big.blow5", "-K", K, "-B", "100G")
        assert [ln.split("\t")[0] for ln in lines] == big_ids
        assert [int(ln.split("\t")[1]) for ln in lines] == [len(s) for s in big_sigs]
        m = re.search(r"total entries: (\d+)", err)
        assert m and int(m.group(1)) == n
