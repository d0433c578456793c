#!/usr/bin/env python
"""Device time of the BLOW5 record decode (sf_inflate_kernel + sf_signal_kernel): the event stage of a batch submitted
as records minus the event stage of the same batch submitted as samples.  Usage: python tools/decode_bench.py [--reads N]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from sigfish_b200 import capi, synth  # noqa: E402
from test_gpu_blow5 import make_records  # noqa: E402


def main():
    n = int(sys.argv[sys.argv.index("--reads") + 1]) if "--reads" in sys.argv else 65536
    k = 6
    lm, _ = synth.make_model(k)
    seq = synth.random_sequence(29903, np.random.default_rng(2))
    uniq, _ = synth.simulate_reads([seq], k, lm, 4096, seed=5, bases_per_read=450)
    ids = [f"read_{i:06d}" for i in range(len(uniq))]
    scs = [synth.DNA_SCALING] * len(uniq)
    recs, spos, sbytes = make_records(ids, uniq, scs, True, True)
    rep = lambda a: [a[i % len(uniq)] for i in range(n)]
    ctx = capi.Context(lm, k)
    ctx.set_ref([seq])
    sigs = rep(uniq)
    ctx.submit(0, *ctx.pack(sigs, rep(scs)))
    ctx.collect(0)
    t_s = ctx.timing(0)
    ev_r = []
    for _ in range(3):
        ctx.submit_records(1, rep(recs), 1, 1, rep(spos), rep(sbytes), [len(s) for s in sigs], rep(scs))
        ctx.collect(1)
        ev_r.append(ctx.timing(1).events_ms)
    print(f"{n} records ({sum(len(r) for r in recs) / len(recs):.0f} B compressed, {np.mean([len(s) for s in uniq]):.0f} samples each): "
          f"events stage from samples {t_s.events_ms:.2f} ms, from records {min(ev_r):.2f} ms -> decode {min(ev_r) - t_s.events_ms:.2f} ms "
          f"= {(min(ev_r) - t_s.events_ms) * 1e3 / n:.3f} us per record; DTW stage {t_s.dtw_ms:.1f} ms")


if __name__ == "__main__":
    main()
