#!/usr/bin/env python
"""bench.py -- sDTW GCUPS (+ reads/s) of the `sigfish dtw` hot path on B200.

Headline workload (BASELINE.json configs[3], "C4"): synthetic R10 DNA (k=9 model), reads of ~450 bases
mapped with the defaults (-q 250 -p 50) against ONE synthetic 1 Mb contig, both strands
(2 x 999 992 reference columns => 5.0e8 DTW cells per read).  A step = one pass of the whole hot
path (event detection + query normalisation + sDTW + hit selection + start coordinate) over one
batch of `--reads` reads per GPU.  Reads shard across GPUs with the reference replicated, so
scaling is weak and there is no collective on the data path.

  value   : whole-job GCUPS, inputs already resident in HBM (sfgpu_resubmit), CUDA-event time on the
            library's stream, max over ranks
  e2e     : the same metric through the C-ABI with HOST buffers (sfgpu_submit + sfgpu_collect):
            host packing, pinned H2D of the int16 signals and D2H of the per-read hits inside the
            timed region
  roofline: the dominant kernel (sf_dtw_pair_kernel) against the fp32 issue-slot floor of the recurrence
            itself (2 FADD + 1 half-rate FMNMX3 = 4 issue slots per cell); the figure against
            the kernel's own instruction stream is kept as a secondary key
  shapes  : the other named shapes of BASELINE.json with the same measurements -- C2 (R9 DNA against a 30 kb
            genome, both strands) and C5 (RNA004 --rna --invert against a 50 000-transcript transcriptome)
  e2e_files: the product as a user runs it: rank 0 starts `sigfish-b200 dtw ref.fa reads.blow5 --gpus N` on
            files of the C4 set (wall clock from exec to exit, start-up included)
  cpu_baseline / --impl reference: the unmodified reference binary (oracle/_ref/sigfish, built from
            /root/reference by oracle/Makefile) with -t <host cores> on a bounded sample of the
            same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

OUT = sys.stdout
REF_LEN = 1_000_000
F_RNA, F_INV = 0x001, 0x004

# name -> workload.  reads: reads per GPU per step (0: four full waves of DTW tasks); unique: distinct reads
# simulated (the batch repeats them), because simulating 65 k reads in numpy would take longer than measuring them
SHAPES = {
    "C4": dict(workload="C4: synthetic R10 DNA (k=9), reads x q=250 vs one 1 Mb contig, both strands",
               k=9, flags=0, q=250, p=50, kit="sqk-lsk114", reads=0, unique=0),
    "C2": dict(workload="C2: synthetic R9 DNA (k=6), reads x q=250 vs one 29 903-base genome (nCoV size), both strands",
               k=6, flags=0, q=250, p=50, kit="sqk-lsk109", reads=65536, unique=8192),
    "C5": dict(workload="C5: synthetic RNA004 (k=9) --rna --invert, reads x q=250 vs 50 000 transcripts of 400-4000 nt "
                        "(375 columns each)",
               k=9, flags=F_RNA | F_INV, q=250, p=50, kit="sqk-rna004", reads=2048, unique=2048),
}

# q = 250 reads run two per warp (sf_dtw_pair_kernel<16,false,9>: 16 lanes x R=16 rows each).  Its inner loop issues
# 104 SASS instructions per warp for one macro-step of 32 lanes x 16 rows x 2 columns (cuobjdump: 64 FADD + 32 FMNMX3 +
# LDS.64 + 2 SHFL + 2 IMAD + STS.64 + loop), i.e. 52 per column of 512 cells.  FMNMX3 runs on the half-rate ALU pipe
# and blocks the issue port for 2 cycles (measured: tools/ubench_alu.cu, profiles/r01_ubench_alu.txt), so a column
# costs 52 + 16 = 68 issue slots.  The floor of the recurrence itself is 2 FADD + 1 FMNMX3 = 4 slots per cell
# (64 per column).  DESIGN.md 5.1.
SASS_PER_STEP = 52.0
ISSUE_SLOTS_PER_STEP = 68.0
ROWS_PER_LANE = 16
FLOOR_SLOTS_PER_CELL = 4.0
# DRAM traffic of one DTW launch from the committed `ncu --set full` capture (dram__bytes_read.sum +
# dram__bytes_write.sum at 8288 reads); almost all of it is wavefront checkpoints.  NOT measured in this run.
NCU_DTW_TRAFFIC = {"reads": 8288, "bytes": 357.556992e6 + 5078.331e6, "source": "profiles/r02_ncu_summary.md"}


def make_inputs(shape: str, n_unique: int, seed: int, ref_len: int = REF_LEN):
    """model, reference sequences (names, seqs), `n_unique` simulated reads and their scaling"""
    from sigfish_b200 import synth
    sp = SHAPES[shape]
    k = sp["k"]
    mean, stdv = synth.make_model(k)
    if shape == "C4":
        names, seqs = ["chrS"], [synth.random_sequence(ref_len, np.random.default_rng(1))]
        sigs, _ = synth.simulate_reads(seqs, k, mean, n_unique, seed=seed, bases_per_read=450)
        sc = synth.DNA_SCALING
    elif shape == "C2":
        names, seqs = ["genome30k"], [synth.random_sequence(29_903, np.random.default_rng(2))]
        sigs, _ = synth.simulate_reads(seqs, k, mean, n_unique, seed=seed, bases_per_read=450)
        sc = synth.DNA_SCALING
    else:
        names, seqs = synth.transcriptome(50_000, 20241, 400, 4000)
        src = [seqs[i] for i in range(7, len(seqs), 97)]  # reads come from 515 of the transcripts (3' ends)
        sigs, _ = synth.simulate_reads(src, k, mean, n_unique, seed=seed, rna=True, bases_per_read=420)
        sc = synth.RNA_SCALING
    return mean, stdv, names, seqs, sigs, sc


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks / throttle reasons while the timed region runs"""

    def __init__(self, device: int):
        super().__init__(daemon=True)
        self.device = device
        self.stop_flag = False
        self.sm, self.reasons, self.sm_max = [], set(), None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.device)], capture_output=True, text=True, timeout=5).stdout.strip()
                f = [x.strip() for x in out.split(",")]
                self.sm.append(float(f[0]))
                self.sm_max = float(f[1])
                for nm, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))
    return {}


# ---------------------------------------------------------------------------------- reference (CPU) arm

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "sigfish")


def host_threads(cols_per_matrix: int, q: int) -> int:
    """the reference allocates a 4*q*rlen byte cost matrix per worker thread (sigfish.c:873)"""
    n = os.cpu_count() or 1
    try:
        avail = int(re.search(r"MemAvailable:\s+(\d+)", open("/proc/meminfo").read()).group(1)) * 1024
        per = 4 * q * cols_per_matrix * 1.25
        n = max(1, min(n, int(avail * 0.7 // per)))
    except Exception:
        pass
    return n


def run_reference_once(workdir: str, shape: str, n_reads: int, threads: int):
    """one run of the unmodified reference binary on the sample; returns (process_db seconds, PAF rows)"""
    sp = SHAPES[shape]
    cmd = [REF_BIN, "dtw", os.path.join(workdir, "ref.fa"), os.path.join(workdir, "reads.slow5"), "--kmer-model",
           os.path.join(workdir, "model.txt"), "-t", str(threads), "-K", str(max(n_reads, 1)), "-q", str(sp["q"]),
           "-p", str(sp["p"]), "-o", os.path.join(workdir, "cpu.paf")]
    if sp["flags"] & F_RNA:
        cmd.append("--rna")
    if sp["flags"] & F_INV:
        cmd.append("--invert")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference binary failed: " + r.stderr[-1000:])
    t = float(re.search(r"Data processing time: ([0-9.]+) sec", r.stderr).group(1))
    rows = sum(1 for _ in open(os.path.join(workdir, "cpu.paf")))
    return t, rows


def prepare_reference_sample(shape: str, n_reads: int, ref_len: int = REF_LEN):
    from sigfish_b200 import synth
    sp = SHAPES[shape]
    d = os.path.join(synth.tmpdir(), "bench_ref_" + shape)
    os.makedirs(d, exist_ok=True)
    mean, stdv, names, seqs, sigs, sc = make_inputs(shape, n_reads, seed=1234, ref_len=ref_len)
    synth.write_model_file(os.path.join(d, "model.txt"), sp["k"], mean, stdv)
    synth.write_fasta(os.path.join(d, "ref.fa"), names, seqs)
    synth.write_slow5_ascii(os.path.join(d, "reads.slow5"), [f"read{i}" for i in range(n_reads)], sigs,
                            rna=bool(sp["flags"] & F_RNA), kit=sp["kit"], scaling=sc)
    return d


def shape_geometry(shape: str, ref_len: int = REF_LEN):
    """(reference columns one read is aligned against, columns of the largest single cost matrix)"""
    sp = SHAPES[shape]
    if shape == "C4":
        n = ref_len + 1 - sp["k"]
        return 2 * n, n
    if shape == "C2":
        n = 29_903 + 1 - sp["k"]
        return 2 * n, n
    return 50_000 * 375, 375


def cpu_sample(shape: str, reads_per_thread: int, runs: int, ref_len: int = REF_LEN):
    """times the reference's stock CPU path on a bounded sample of the shape; returns the cpu_baseline object"""
    sp = SHAPES[shape]
    cols, per_matrix = shape_geometry(shape, ref_len)
    threads = host_threads(per_matrix, sp["q"])
    n = max(threads * reads_per_thread, 8)
    d = prepare_reference_sample(shape, n, ref_len)
    secs = [run_reference_once(d, shape, n, threads)[0] for _ in range(runs)]
    sec = float(np.median(secs))
    return {"value": n * sp["q"] * cols / sec / 1e9, "unit": "GCUPS", "cores": threads, "kind": "reference",
            "reads_per_s": n / sec,
            "sample": f"{n} reads of the workload ({reads_per_thread} per thread, stock path: one 4*q*rlen-byte cost matrix "
                      f"malloc'd and first-touched per read and strand), -t {threads}, median process_db wall of {runs} run(s): "
                      f"{sec:.2f} s"}, sec, n, threads


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not os.path.exists(REF_BIN):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/sigfish was not built (no /root/reference at build time)"}), file=OUT, flush=True)
        return
    sp = SHAPES["C4"]
    cols, per_matrix = shape_geometry("C4", args.ref_len)
    threads = host_threads(per_matrix, sp["q"])
    n_reads = args.cpu_reads or max(2 * threads, 8)
    d = prepare_reference_sample("C4", n_reads, args.ref_len)
    cells_per_read = sp["q"] * cols
    times = []
    for i in range(args.warmup + args.steps):
        t, rows = run_reference_once(d, "C4", n_reads, threads)
        if i >= args.warmup:
            times.append(t)
    sec = float(np.mean(times))
    gcups = n_reads * cells_per_read / sec / 1e9
    sample = f"{n_reads} reads of the workload per step ({n_reads * cells_per_read:.3g} cells), -t {threads}, process_db wall time"
    line = {"impl": "reference", "metric": "sDTW GCUPS", "value": gcups, "unit": "GCUPS", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "reads_per_s": n_reads / sec,
            "config": {"workload": sp["workload"], "reads_per_step": n_reads, "query_size": sp["q"], "prefix_size": sp["p"],
                       "ref_columns": cols, "host": "unmodified reference binary, CPU only"},
            "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=OUT, flush=True)


# ---------------------------------------------------------------------------------- our arm

def run_shape(shape, args, R, rank, local, flush, with_clocks, steps=None):
    import torch
    from sigfish_b200 import capi, synth
    sp = SHAPES[shape]
    steps = steps or args.steps
    n_reads = args.reads if (shape == "C4" and args.reads > 0) else sp["reads"]
    t_prep = time.perf_counter()
    ref_len = args.ref_len if shape == "C4" else REF_LEN
    mean, stdv = synth.make_model(sp["k"])
    # the context first: the default batch of the headline shape is sized from the GPU (four waves of DTW tasks)
    ctx = capi.Context(mean, sp["k"], flags=sp["flags"], query_size=sp["q"], prefix_size=sp["p"], device=local, n_slots=2,
                       warm_blocks=args.warm_blocks, piece_periods=args.piece_periods,
                       pore=2 if sp["kit"] == "sqk-rna004" else (1 if sp["kit"] == "sqk-lsk114" else 0))
    n_unique = sp["unique"]
    _, _, names, seqs, _, sc1 = make_inputs(shape, 0, seed=0, ref_len=ref_len)
    ctx.set_ref(seqs)
    if n_reads <= 0:
        n_reads = 4 * ctx.wave_reads
    if n_unique <= 0 or n_unique > n_reads:
        n_unique = n_reads
    _, _, _, _, uniq, _ = make_inputs(shape, n_unique, seed=100 + rank, ref_len=ref_len)
    sigs = [uniq[i % n_unique] for i in range(n_reads)]
    sc = [sc1] * len(sigs)
    packed = ctx.pack(sigs, sc)
    ref_cols = ctx.ref_columns
    prep_s = time.perf_counter() - t_prep

    # ---- device-resident throughput ----
    ctx.submit(0, *packed)
    first = ctx.collect(0).copy()
    for _ in range(args.warmup):
        ctx.resubmit(0)
        ctx.collect(0)
    sampler = ClockSampler(local)
    if with_clocks and rank == 0:  # one nvidia-smi poller per job, not per rank
        sampler.start()
    R.barrier()
    tot = dtw = evt = trc = 0.0
    cells = 0.0
    launches = 0
    split = {}
    t_wall0 = time.perf_counter()
    for _ in range(steps):
        flush.zero_()                 # L2 flush between timed iterations (outside the device-timed region)
        torch.cuda.synchronize()
        ctx.resubmit(0)
        t = ctx.timing(0)             # waits for the step; CUDA events on the library's stream
        tot += t.events_ms + t.dtw_ms + t.trace_ms
        dtw += t.dtw_ms
        evt += t.events_ms
        trc += t.trace_ms
        cells = t.cells
        launches += t.dtw_launches + t.other_launches
        split = {"tasks_per_read": t.tasks_per_read, "piece_columns": t.piece_blocks * 64, "redone_pieces_last_step": t.redone_pieces}
    R.barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    sampler.stop_flag = True
    if with_clocks and rank == 0:
        sampler.join()
    last = ctx.collect(0)
    assert last.tobytes() == first.tobytes(), "results changed between steps"
    mapped = int((last["qlen"] > 0).sum())
    ms_step = tot / steps

    # ---- end to end through the C-ABI with host buffers (double-buffered slots) ----
    # untimed warm-up of the second slot: its pinned and device buffers are allocated on first use
    ctx.submit(1, *packed)
    ctx.collect(1)
    R.barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        ctx.submit(i & 1, *packed)
        if i > 0:
            ctx.collect((i - 1) & 1)
    ctx.collect((steps - 1) & 1)
    R.barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / steps
    h2d = int(packed[0].nbytes + (2 * len(sigs) + 1) * 8 + 3 * 4 * len(sigs))
    d2h = int(len(sigs) * (40 + 32))
    ctx.close()

    # slowest rank sets the step time; work is summed over ranks
    ms_step, e2e_ms, dtw_ms, evt_ms, trc_ms, wall_step = R.max(
        [ms_step, e2e_ms, dtw / steps, evt / steps, trc / steps, wall_ms / steps])
    job_cells, job_reads, job_samples = R.sum([cells, float(len(sigs)), float(sum(len(s) for s in sigs))])
    return dict(shape=shape, n_reads=len(sigs), n_unique=n_unique, ref_cols=int(ref_cols), cells=cells, mapped=mapped,
                ms_step=ms_step, e2e_ms=e2e_ms, dtw_ms=dtw_ms, evt_ms=evt_ms, trc_ms=trc_ms, wall_step=wall_step,
                job_cells=job_cells, job_reads=job_reads, job_samples=job_samples, launches=launches, split=split,
                h2d=h2d, d2h=d2h, clocks=sampler.result() if (with_clocks and rank == 0) else None, prep_s=prep_s,
                samples=int(sum(len(s) for s in sigs)), steps=steps)


def e2e_from_files(args, world):
    """rank 0: `sigfish-b200 dtw` on FASTA / BLOW5 / model files of the C4 set with --gpus <world>; wall clock of the
    whole command (process start-up, CUDA contexts, reference synthesis, BLOW5 decoding, output) and its own timers"""
    from sigfish_b200 import build as B
    from sigfish_b200 import synth
    sp = SHAPES["C4"]
    d = os.path.join(synth.tmpdir(), "bench_files")
    os.makedirs(d, exist_ok=True)
    n_unique = 8288
    mean, stdv, names, seqs, sigs, sc = make_inputs("C4", n_unique, seed=4242, ref_len=args.ref_len)
    synth.write_model_file(os.path.join(d, "model.txt"), sp["k"], mean, stdv)
    synth.write_fasta(os.path.join(d, "ref.fa"), names, seqs)
    one = os.path.join(d, "unit.blow5")
    synth.write_blow5(one, [f"read_{i:06d}" for i in range(n_unique)], sigs, kit=sp["kit"])
    # the file of the run: the unit's records repeated (header once, end marker once)
    raw = open(one, "rb").read()
    hlen = 64 + 4 + int.from_bytes(raw[64:68], "little")
    body = raw[hlen:-5]
    reps = max(1, round(args.files_reads_per_gpu * world / n_unique))
    path = os.path.join(d, "reads.blow5")
    with open(path, "wb") as f:
        f.write(raw[:hlen])
        for _ in range(reps):
            f.write(body)
        f.write(b"5WOLB")
    n_reads = reps * n_unique
    cmd = [B.CLI, "dtw", os.path.join(d, "ref.fa"), path, "--kmer-model", os.path.join(d, "model.txt"), "-t",
           str(os.cpu_count() or 8), "--gpus", str(world), "-o", os.path.join(d, "out.paf"), "--verbose", "5"]
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, SFGPU_TRACE="1"))
    wall = time.perf_counter() - t0
    log_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(log_dir):  # the command's own timeline (--verbose 5, SFGPU_TRACE), kept beside the bench line
        open(os.path.join(log_dir, f"bench_cli_stderr_{world}gpu.txt"), "w").write(r.stderr)
    if r.returncode != 0:
        return {"error": r.stderr[-500:]}
    out = {"reads": n_reads, "gpus": world, "wall_s": wall, "reads_per_s_wall": n_reads / wall,
           "GCUPS_wall": n_reads * sp["q"] * shape_geometry("C4", args.ref_len)[0] / wall / 1e9,
           "paf_rows": sum(1 for _ in open(os.path.join(d, "out.paf")))}
    for line in r.stderr.splitlines():
        for key, name in (("Data loading time", "load_s"), ("Data processing time", "processing_s"), ("Parse time", "parse_s"),
                          ("DTW time", "dtw_gpu0_s"), ("Events + normalise", "events_gpu0_s"), ("Data output time", "output_s")):
            if key in line:
                out[name] = float(line.split(":")[-1].split()[0])
        m = re.match(r"\[init_core::([0-9.]+)\] GPU contexts", line)
        if m:
            out["init_s"] = float(m.group(1))
        m = re.match(r"\[dtw_main::([0-9.]+)\*", line)
        if m and "first_batch_loaded_s" not in out and "loaded" in line:
            out["first_batch_loaded_s"] = float(m.group(1))
    if "processing_s" in out:
        out["reads_per_s_processing"] = n_reads / out["processing_s"]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=0,
                    help="reads per GPU per step of the headline shape (0: four full waves of DTW tasks, 4 x sfgpu_wave_reads() = 16576 on a 148-SM B200)")
    ap.add_argument("--ref-len", type=int, default=REF_LEN)
    ap.add_argument("--cpu-reads", type=int, default=0, help="reads in the CPU baseline sample (0: two per host thread)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--shapes", default="C2,C5", help="further shapes to measure (comma separated, '' for none)")
    ap.add_argument("--no-files", action="store_true", help="skip the from-files run of the command line")
    ap.add_argument("--files-reads-per-gpu", type=int, default=16576)
    ap.add_argument("--warm-blocks", type=int, default=0, help="experiment knob: warm-up of a piece in 64-column blocks (0: 2q columns)")
    ap.add_argument("--piece-periods", type=int, default=0, help="experiment knob: piece length in checkpoint periods (0: per batch, <0: no splitting)")
    args = ap.parse_args()
    # the JSON line is the only thing that may reach stdout: libraries that write to fd 1 (NCCL prints its version
    # there) are sent to stderr, and the line itself goes to the saved descriptor
    global OUT
    sys.stdout.flush()
    OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = max(args.warmup, 1)

    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    from sigfish_b200 import ranks

    rank, world, local = ranks.env_rank()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    R = ranks.Ranks(backend="nccl", device=torch.device("cuda", local))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    m = run_shape("C4", args, R, rank, local, flush, with_clocks=True)
    extra = [s for s in args.shapes.split(",") if s and s in SHAPES and s != "C4"]
    # the further shapes are measured over at most 5 steps each (the contract's K steps apply to the headline)
    others = {s: run_shape(s, args, R, rank, local, flush, with_clocks=False, steps=min(args.steps, 5)) for s in extra}
    del flush
    torch.cuda.empty_cache()
    files = None
    if not args.no_files:
        # the other ranks wait on the CPU: an NCCL barrier would keep a kernel spinning on their GPUs, which the
        # command line started by rank 0 is about to use
        R.barrier()
        R.host_barrier()
        if rank == 0:
            try:
                files = e2e_from_files(args, world)
            except Exception as e:  # reported, never allowed to sink the bench line
                files = {"error": str(e)}
        R.host_barrier()

    if rank == 0:
        clocks = m["clocks"]
        peaks = measured_peaks()
        sm_count = torch.cuda.get_device_properties(local).multi_processor_count
        clk = (clocks["sm_mhz"] or peaks.get("sm_max_mhz", 1965.0)) * 1e6
        lane_slots = sm_count * 4 * clk * 32           # lane-instruction issue slots per second
        floor_peak = lane_slots / FLOOR_SLOTS_PER_CELL  # cells/s if nothing but the recurrence were issued
        own_peak = sm_count * 4 * clk / ISSUE_SLOTS_PER_STEP * (32 * ROWS_PER_LANE)
        naive_peak = sm_count * 4 * clk / SASS_PER_STEP * (32 * ROWS_PER_LANE)
        sp = SHAPES["C4"]
        dtw_cells_per_s = m["cells"] / (m["dtw_ms"] * 1e-3)
        value = ranks.job_throughput(m["job_cells"], m["ms_step"])
        line = {
            "metric": "sDTW GCUPS", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": m["ms_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "reads_per_s": m["job_reads"] / (m["ms_step"] * 1e-3),
            "config": {"workload": sp["workload"], "reads_per_step_per_gpu": m["n_reads"], "query_size": sp["q"], "prefix_size": sp["p"],
                       "kmer": sp["k"], "ref_columns": m["ref_cols"], "cells_per_step_per_gpu": m["cells"],
                       "samples_per_step_per_gpu": m["samples"], "mapped_reads": m["mapped"],
                       "batch": "four full waves of (read, strand) DTW tasks per step (4 x sfgpu_wave_reads)" if args.reads <= 0 else "--reads",
                       "l2": "256 MB buffer written between timed iterations (L2 flush)",
                       "parallelism": f"reads sharded over {world} GPU(s), reference replicated, no collective"},
            "split": m["split"],
            "stage_ms": {"events": m["evt_ms"], "dtw": m["dtw_ms"], "merge_trace": m["trc_ms"], "wall_per_step_incl_flush": m["wall_step"]},
            "e2e": {"value": m["job_cells"] / (m["e2e_ms"] * 1e-3) / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": m["h2d"],
                    "d2h_bytes_per_step": m["d2h"], "ms_per_step": m["e2e_ms"], "reads_per_s": m["job_reads"] / (m["e2e_ms"] * 1e-3)},
            "gpu_launches": m["launches"],
            "clocks": clocks,
            "roofline": {"bound": "alu-issue", "kernel": "sf_dtw_pair_kernel<16,false,9>",
                         "achieved": dtw_cells_per_s / 1e9, "peak": floor_peak / 1e9, "unit": "GCUPS",
                         "frac": dtw_cells_per_s / floor_peak,
                         "peak_source": f"recurrence floor: {sm_count} SMs x 4 schedulers x 32 lanes x {clk / 1e6:.0f} MHz (median under load) / "
                                        f"{FLOOR_SLOTS_PER_CELL:.0f} issue slots per cell (FADD + half-rate FMNMX3 + FADD; "
                                        "half rate measured, profiles/r01_ubench_alu.txt)",
                         "traffic": NCU_DTW_TRAFFIC["bytes"] * m["n_reads"] / NCU_DTW_TRAFFIC["reads"],
                         "traffic_measured_in_this_run": False,
                         "traffic_source": NCU_DTW_TRAFFIC["source"] + " (ncu --set full at 8288 reads, scaled by reads); "
                                           "algorithmic HBM bytes are 0.016 B/cell (8 MB reference stream, L2 resident) "
                                           "plus the wavefront checkpoints and the pieces' warm fronts (0.66 MB/read)",
                         "frac_of_own_instruction_stream": dtw_cells_per_s / own_peak,
                         "own_instruction_stream": f"{ISSUE_SLOTS_PER_STEP:.0f} issue slots per 32x{ROWS_PER_LANE} cells "
                                                   f"({SASS_PER_STEP:.0f} SASS per column, the {ROWS_PER_LANE} half-rate FMNMX3 counted twice)",
                         "frac_if_every_sass_were_one_slot": dtw_cells_per_s / naive_peak,
                         "events_kernel_GBps": (m["job_samples"] / world) * 2 / (m["evt_ms"] * 1e-3) / 1e9 if m["evt_ms"] > 0 else None,
                         "hbm_peak_GBps": peaks.get("hbm_gbs")},
        }
        shapes = {}
        for s, o in others.items():
            dcs = o["cells"] / (o["dtw_ms"] * 1e-3)
            shapes[s] = {
                "workload": SHAPES[s]["workload"], "value": ranks.job_throughput(o["job_cells"], o["ms_step"]), "unit": "GCUPS",
                "reads_per_s": o["job_reads"] / (o["ms_step"] * 1e-3), "ms_per_step": o["ms_step"], "steps": o["steps"],
                "reads_per_step_per_gpu": o["n_reads"], "distinct_reads": o["n_unique"], "ref_columns": o["ref_cols"],
                "mapped_reads": o["mapped"],
                "stage_ms": {"events": o["evt_ms"], "dtw": o["dtw_ms"], "merge_trace": o["trc_ms"]},
                "split": o["split"],
                "e2e": {"value": o["job_cells"] / (o["e2e_ms"] * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": o["e2e_ms"],
                        "reads_per_s": o["job_reads"] / (o["e2e_ms"] * 1e-3), "h2d_bytes_per_step": o["h2d"], "d2h_bytes_per_step": o["d2h"]},
                "dtw_frac_of_recurrence_floor": dcs / floor_peak,
                "step_frac_of_recurrence_floor": (o["cells"] / (o["ms_step"] * 1e-3)) / floor_peak,
            }
        if shapes:
            line["shapes"] = shapes
        if files is not None:
            line["e2e_files"] = files
        if world == 1 and not args.no_cpu_baseline:
            for s in ["C4"] + list(shapes):
                try:
                    if not os.path.exists(REF_BIN):
                        raise RuntimeError("oracle/_ref/sigfish missing")
                    if s == "C4" and args.cpu_reads:
                        cols, per_matrix = shape_geometry("C4", args.ref_len)
                        threads = host_threads(per_matrix, sp["q"])
                        d = prepare_reference_sample("C4", args.cpu_reads, args.ref_len)
                        sec = float(np.median([run_reference_once(d, "C4", args.cpu_reads, threads)[0] for _ in range(3)]))
                        cb = {"value": args.cpu_reads * sp["q"] * cols / sec / 1e9, "unit": "GCUPS", "cores": threads,
                              "kind": "reference", "reads_per_s": args.cpu_reads / sec,
                              "sample": f"{args.cpu_reads} reads of the workload, median of 3 runs, -t {threads}, process_db wall {sec:.2f} s"}
                    else:
                        # C4 / C2: two reads per thread, three runs; C5 (4.7e9 cells per read): one read per thread, one run
                        cb = cpu_sample(s, 1 if s == "C5" else 2, 1 if s == "C5" else 3, args.ref_len)[0]
                except Exception as e:  # the baseline is reported, never allowed to sink the bench line
                    cb = {"value": None, "unit": "GCUPS", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}
                if s == "C4":
                    line["cpu_baseline"] = cb
                else:
                    shapes[s]["cpu_baseline"] = cb
        print(json.dumps(line), file=OUT, flush=True)
    R.close()


if __name__ == "__main__":
    main()
