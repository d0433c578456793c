#!/usr/bin/env python
"""End-to-end, from files: wall time and reads/s of `sigfish-b200 dtw` on FASTA / BLOW5 / model files of the
C4 shape (one 1 Mb contig, R10 k=9 reads).  Development / evidence tool: the judged numbers come from bench.py;
the byte-for-byte comparison with the reference binary at this scale is tests/test_gpu_cli.py
(test_cli_c4_scale_matches_reference_binary).
Usage: python tools/cli_e2e.py [--reads N] [--gpus G] [-K N]"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sigfish_b200 import build as B  # noqa: E402
from sigfish_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=5920)
    ap.add_argument("--ref-len", type=int, default=1_000_000)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("-K", type=int, default=0)
    ap.add_argument("--shape", default="c4", choices=["c4", "c2"], help="c4: R10 k=9 reads vs --ref-len contig; c2: R9 k=6 reads vs a 29 903-base genome")
    ap.add_argument("--auto-batch", action="store_true", help="leave -K / -B to the command line's own choice")
    ap.add_argument("--trace", default="", help="write the CLI's stderr (--verbose 5, SFGPU_TRACE=1) to this file")
    args = ap.parse_args()
    B.build_all()
    d = os.path.join(synth.tmpdir(), "cli_e2e")
    os.makedirs(d, exist_ok=True)
    k = 9 if args.shape == "c4" else 6
    if args.shape == "c2":
        args.ref_len = 29_903
    mean, stdv = synth.make_model(k)
    seq = synth.random_sequence(args.ref_len, np.random.default_rng(1))
    # a unit of at most 8288 distinct reads, its records repeated up to --reads (writing BLOW5 from Python costs
    # 0.5 ms per read, which on a multi-GPU box is paid per GPU-minute)
    unit = min(args.reads, 8288)
    sigs, _ = synth.simulate_reads([seq], k, mean, unit, seed=4242, bases_per_read=450)
    ids = [f"read_{i:06d}" for i in range(len(sigs))]
    synth.write_model_file(os.path.join(d, "model.txt"), k, mean, stdv)
    synth.write_fasta(os.path.join(d, "ref.fa"), ["chrS"], [seq])
    one = os.path.join(d, "unit.blow5")
    synth.write_blow5(one, ids, sigs, kit="sqk-lsk114" if k == 9 else "sqk-lsk109")
    raw = open(one, "rb").read()
    hlen = 64 + 4 + int.from_bytes(raw[64:68], "little")
    reps = max(1, round(args.reads / unit))
    args.reads = reps * unit
    with open(os.path.join(d, "reads.blow5"), "wb") as f:
        f.write(raw[:hlen])
        for _ in range(reps):
            f.write(raw[hlen:-5])
        f.write(b"5WOLB")
    K = args.K or args.reads
    out = {"reads": args.reads, "gpus": args.gpus, "cells_per_read": 250 * 2 * (args.ref_len + 1 - k)}

    t0 = time.perf_counter()
    r = subprocess.run([B.CLI, "dtw", os.path.join(d, "ref.fa"), os.path.join(d, "reads.blow5"), "--kmer-model",
                        os.path.join(d, "model.txt")] + ([] if args.auto_batch else ["-K", str(K), "-B", "100G"]) +
                       ["-t", str(os.cpu_count() or 8), "--gpus", str(args.gpus),
                        "-o", os.path.join(d, "gpu.paf")] + (["--verbose", "5"] if args.trace else []),
                       capture_output=True, text=True, env=dict(os.environ, **({"SFGPU_TRACE": "1"} if args.trace else {})))
    if args.trace:
        open(args.trace, "w").write(r.stderr)
    out["b200_wall_s"] = time.perf_counter() - t0
    assert r.returncode == 0, r.stderr[-3000:]
    for line in r.stderr.splitlines():
        for key in ("Data loading time", "Data processing time", "Parse time", "DTW time", "Events + normalise"):
            if key in line:
                out["b200 " + key] = float(line.split(":")[-1].split()[0])
    out["b200_reads_per_s_processing"] = args.reads / out["b200 Data processing time"]
    out["b200_GCUPS_processing"] = args.reads * out["cells_per_read"] / out["b200 Data processing time"] / 1e9

    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
