#!/usr/bin/env python
"""Host-feed ceiling of `sigfish-b200 dtw`: the CLI's host code (sigfish_b200/host/*.c) linked against a NULL
device (tools/hostfeed/sfgpu_null.c) and run on a synthetic BLOW5 file -- no GPU involved, nothing is aligned.
Prints the read rate at which BLOW5 load + record decode + sharding + staging copy + PAF epilogue + output
saturate on this machine's cores, i.e. how many reads/s the host can feed to the GPUs.
Usage: python tools/hostfeed/run.py [--reads N] [--gpus G] [-t T] [-K K]"""
import argparse
import glob
import json
import os
import re
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from sigfish_b200 import synth  # noqa: E402

BUILD = os.path.join(HERE, "_build")


def build():
    os.makedirs(BUILD, exist_ok=True)
    inc = os.path.join(ROOT, "include")
    host = os.path.join(ROOT, "sigfish_b200", "host")
    subprocess.run(["gcc", "-O2", "-std=c99", "-Wall", "-fPIC", "-shared", "-I", inc, "-o", os.path.join(BUILD, "libsfgpu.so"),
                    os.path.join(HERE, "sfgpu_null.c")], check=True)
    srcs = sorted(glob.glob(os.path.join(host, "*.c")))
    exe = os.path.join(BUILD, "sigfish-hostfeed")
    subprocess.run(["gcc", "-O2", "-g", "-std=c99", "-Wall", "-D_GNU_SOURCE", "-I", inc, "-o", exe] + srcs +
                   ["-L", BUILD, "-lsfgpu", "-Wl,-rpath," + BUILD, "-lz", "-lpthread", "-lm"], check=True)
    return exe


def one_run(cmd, env, args, blow5):
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    wall = time.perf_counter() - t0
    assert r.returncode == 0, r.stderr[-3000:]
    out = {"reads": args.reads, "gpus": args.gpus, "threads": args.t, "cores": os.cpu_count(), "wall_s": wall,
           "blow5_MB": os.path.getsize(blow5) / 1e6}
    for line in r.stderr.splitlines():
        for key in ("Data loading time", "Data processing time", "Parse time", "Data output time", "total entries"):
            if key in line:
                try:
                    out[key] = float(line.split(":")[-1].split()[0])
                except ValueError:
                    pass
    # duration of the batch loop from the progress lines: "[dtw_main::<seconds>*<cpu>] N Entries (...) loaded|processed"
    loaded = [float(m.group(1)) for m in re.finditer(r"\[dtw_main::([0-9.]+)\*[0-9.]+\] \d+ Entries .* loaded", r.stderr)]
    done = [float(m.group(1)) for m in re.finditer(r"\[dtw_main::([0-9.]+)\*[0-9.]+\] \d+ Entries .* processed", r.stderr)]
    if loaded and done:
        first_load = out.get("Data loading time", 0.0) / len(loaded)
        out["batches"] = len(loaded)
        out["loop_s"] = done[-1] - (loaded[0] - first_load)
        out["host_reads_per_s"] = args.reads / out["loop_s"]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=40000)
    ap.add_argument("--gpus", type=int, default=8)
    ap.add_argument("-t", type=int, default=os.cpu_count() or 8)
    ap.add_argument("-K", type=int, default=0)
    ap.add_argument("--repeat", type=int, default=1, help="runs; the median one is reported (shared machines are noisy)")
    args = ap.parse_args()
    exe = build()
    d = os.path.join(synth.tmpdir(), "hostfeed")
    os.makedirs(d, exist_ok=True)
    k = 9
    mean, stdv = synth.make_model(k)
    seq = synth.random_sequence(1_000_000, np.random.default_rng(1))
    blow5 = os.path.join(d, f"reads_{args.reads}.blow5")
    if not os.path.exists(blow5):
        # a unit of at most 4000 distinct reads, its records repeated (compressing every record from Python costs
        # 0.2 ms per read: 300 k reads took over a minute, which once cost a whole gpurun call its time limit)
        unit = min(args.reads, 4000)
        base, _ = synth.simulate_reads([seq], k, mean, unit, seed=4242, bases_per_read=450)
        one = blow5 + ".unit"
        synth.write_blow5(one, [f"read_{i:07d}" for i in range(unit)], base, kit="sqk-lsk114")
        raw = open(one, "rb").read()
        os.remove(one)
        hlen = 64 + 4 + int.from_bytes(raw[64:68], "little")
        reps, rest = divmod(args.reads, unit)
        body = raw[hlen:-5]
        tmp = blow5 + ".tmp"
        with open(tmp, "wb") as f:
            f.write(raw[:hlen])
            for _ in range(reps):
                f.write(body)
            o = 0
            for _ in range(rest):  # whole records: u64 size + bytes
                o += 8 + int.from_bytes(body[o:o + 8], "little")
            f.write(body[:o])
            f.write(b"5WOLB")
        os.replace(tmp, blow5)  # never leave a truncated file behind
    synth.write_model_file(os.path.join(d, "model.txt"), k, mean, stdv)
    synth.write_fasta(os.path.join(d, "ref.fa"), ["chrS"], [seq])
    cmd = [exe, "dtw", os.path.join(d, "ref.fa"), blow5, "--kmer-model", os.path.join(d, "model.txt"), "-t", str(args.t),
           "--gpus", str(args.gpus), "-o", os.path.join(d, "null.paf")]
    if args.K:
        cmd += ["-K", str(args.K), "-B", "100G"]
    env = dict(os.environ, HOSTFEED_GPUS=str(args.gpus))
    runs = [one_run(cmd, env, args, blow5) for _ in range(max(1, args.repeat))]
    runs.sort(key=lambda o: o.get("host_reads_per_s", 0.0))
    out = runs[len(runs) // 2]
    out["all_runs_reads_per_s"] = [round(o.get("host_reads_per_s", 0.0)) for o in runs]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
