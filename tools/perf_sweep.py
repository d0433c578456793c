#!/usr/bin/env python
"""Per-stage device timings of the hot path on the workload shapes of BASELINE.json (development aid; the
judged numbers come from bench.py).  Usage: python tools/perf_sweep.py [c2] [c4] [c5] [--reads N] [--waves N] [-q Q] [--no-pair] [--ragged N] [--min-window COLS] [--only-std]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sigfish_b200 import capi, synth  # noqa: E402


def run(name, ctx, sigs, sc, reps=3):
    packed = ctx.pack(sigs, sc)
    ctx.submit(0, *packed)
    ctx.collect(0)
    rows = []
    for _ in range(reps):
        ctx.resubmit(0)
        t = ctx.timing(0)
        rows.append((t.events_ms, t.dtw_ms, t.trace_ms))
    t0 = time.perf_counter()
    ctx.submit(0, *packed)
    out = ctx.collect(0)
    wall = (time.perf_counter() - t0) * 1e3
    ev, dtw, tr = np.median(np.array(rows), axis=0)
    t = ctx.timing(0)
    print(f"{name:28s} reads={len(sigs):6d} cols={ctx.ref_columns:9d} cells={t.cells:.3g}  events={ev:8.3f} ms  dtw={dtw:9.3f} ms  "
          f"trace={tr:7.3f} ms  h2d={t.h2d_ms:6.3f} d2h={t.d2h_ms:6.3f}  e2e_wall={wall:9.3f} ms  "
          f"dtw_GCUPS={t.cells / dtw / 1e6:8.1f}  total_GCUPS={t.cells / (ev + dtw + tr) / 1e6:8.1f}  "
          f"reads/s={len(sigs) / (ev + dtw + tr) * 1e3:10.0f}  mapped={(out['qlen'] > 0).sum()}", flush=True)


def main():
    which = [a for a in sys.argv[1:] if a in ("c2", "c4", "c5")] or ["c2", "c4", "c5"]
    n_reads = 4096
    if "--reads" in sys.argv:
        n_reads = int(sys.argv[sys.argv.index("--reads") + 1])
    q = 250
    if "-q" in sys.argv:
        q = int(sys.argv[sys.argv.index("-q") + 1])
    nopair = "--no-pair" in sys.argv
    waves = int(sys.argv[sys.argv.index("--waves") + 1]) if "--waves" in sys.argv else 0
    minwin = int(sys.argv[sys.argv.index("--min-window") + 1]) if "--min-window" in sys.argv else 0
    rng = np.random.default_rng(3)
    if "c2" in which:  # R9 DNA vs a 30 kb genome, both strands
        k = 6
        lm, _ = synth.make_model(k)
        seq = synth.random_sequence(29903, rng)
        sigs, _ = synth.simulate_reads([seq], k, lm, n_reads, seed=5, bases_per_read=450)
        ctx = capi.Context(lm, k, query_size=q, no_pairing=nopair, min_window=minwin)
        ctx.set_ref([seq])
        run(f"C2 30kb DNA q250{' nopair' if nopair else ''}", ctx, sigs, [synth.DNA_SCALING] * len(sigs))
        ctx.close()
    if "c4" in which:
        k = 9
        lm, _ = synth.make_model(k)
        seq = synth.random_sequence(1_000_000, rng)
        ctx = capi.Context(lm, k, query_size=q, no_pairing=nopair, min_window=minwin)
        ctx.set_ref([seq])
        if waves:
            n_reads = waves * ctx.wave_reads
        sigs, _ = synth.simulate_reads([seq], k, lm, n_reads, seed=6, bases_per_read=max(450, q + 200))
        if "--ragged" in sys.argv:  # every N-th read cut short: fewer than p+q events, a ragged query
            step = int(sys.argv[sys.argv.index("--ragged") + 1])
            sigs = [s[:len(s) // 3] if i % step == 0 else s for i, s in enumerate(sigs)]
        run(f"C4 1Mb R10 DNA q{q}{' nopair' if nopair else ''}", ctx, sigs, [synth.DNA_SCALING] * len(sigs))
        ctx.close()
    if "c5" in which:  # RNA004-like: many transcripts, 375 columns each, --rna --invert
        k = 9
        lm, _ = synth.make_model(k)
        n_tx = 5000
        seqs = [synth.random_sequence(int(n), rng) for n in rng.integers(400, 4000, size=n_tx)]
        sigs, _ = synth.simulate_reads(seqs, k, lm, min(n_reads, 1024), seed=7, rna=True, bases_per_read=max(420, q + 450))
        variants = ((capi.SFGPU_RNA | capi.SFGPU_INV, "C5 5k transcripts inv"), (capi.SFGPU_RNA, "C5 5k transcripts"),
                    (capi.SFGPU_RNA | capi.SFGPU_DTW, "C5 5k transcripts dtw-std"))
        if "--only-std" in sys.argv:
            variants = variants[2:]
        for flags, nm in variants:
            ctx = capi.Context(lm, k, flags=flags, pore=2, query_size=q, no_pairing=nopair)
            ctx.set_ref(seqs)
            run(nm, ctx, sigs, [synth.RNA_SCALING] * len(sigs))
            ctx.close()


if __name__ == "__main__":
    main()
