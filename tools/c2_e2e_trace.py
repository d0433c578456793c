import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from sigfish_b200 import capi, synth
k = 6
lm, _ = synth.make_model(k)
seq = synth.random_sequence(29903, np.random.default_rng(2))
uniq, _ = synth.simulate_reads([seq], k, lm, 8192, seed=5, bases_per_read=450)
sigs = [uniq[i % 8192] for i in range(65536)]
ctx = capi.Context(lm, k)
ctx.set_ref([seq])
packed = ctx.pack(sigs, [synth.DNA_SCALING] * len(sigs))
ctx.submit(0, *packed); ctx.collect(0); ctx.submit(1, *packed); ctx.collect(1)
print("---- timed", file=sys.stderr, flush=True)
t0 = time.perf_counter()
n = 4
for i in range(n):
    ctx.submit(i & 1, *packed)
    if i > 0:
        ctx.collect((i - 1) & 1)
ctx.collect((n - 1) & 1)
print("per step ms", (time.perf_counter() - t0) * 1e3 / n, file=sys.stderr)
