#!/usr/bin/env python
"""bench.py -- sDTW GCUPS (+ reads/s) of the `sigfish dtw` hot path on B200.

Workload (BASELINE.json configs[3], "C4"): synthetic R10 DNA (k=9 model), reads of ~450 bases
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
  roofline: the dominant kernel (sf_dtw_score_kernel) against the fp32/int ISSUE-SLOT roofline:
            peak cells/s = SMs x 4 schedulers x sustained SM clock / (SASS instructions issued per
            warp per 32 cells); see DESIGN.md
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
KMER = 9
Q, P = 250, 50
WORKLOAD = "C4: synthetic R10 DNA (k=9), reads x q=250 vs one 1 Mb contig, both strands"

# q = 250 reads run two per warp (sf_dtw_pair_kernel<16,false,9>: 16 lanes x R=16 rows each).  Its inner loop issues
# 104 SASS instructions per warp for one macro-step of 32 lanes x 16 rows x 2 columns (cuobjdump: 64 FADD + 32 FMNMX3 +
# LDS.64 + 2 SHFL + 2 IMAD + STS.64 + loop), i.e. 52 per column of 512 cells.  FMNMX3 runs on the half-rate ALU pipe
# and blocks the issue port for 2 cycles (measured: tools/ubench_alu.cu, profiles/r01_ubench_alu.txt), so a column
# costs 52 + 16 = 68 issue slots.  The floor of the recurrence itself is 2 FADD + 1 FMNMX3 = 4 slots per cell
# (64 per column).  DESIGN.md 5.1.
SASS_PER_STEP = 52.0
ISSUE_SLOTS_PER_STEP = 68.0
ROWS_PER_LANE = 16
# DRAM traffic of one DTW launch from the committed `ncu --set full` capture (profiles/r01_ncu_summary.md:
# dram__bytes_read.sum + dram__bytes_write.sum at 8288 reads); almost all of it is wavefront checkpoints
NCU_DTW_TRAFFIC = {"reads": 8288, "bytes": 144.03328e6 + 4778.365e6}


def make_workload(n_reads: int, seed: int, ref_len: int = REF_LEN):
    from sigfish_b200 import synth
    mean, stdv = synth.make_model(KMER)
    rng = np.random.default_rng(1)
    seq = synth.random_sequence(ref_len, rng)
    sigs, _ = synth.simulate_reads([seq], KMER, mean, n_reads, seed=seed, bases_per_read=450)
    return mean, stdv, seq, sigs


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


# ---------------------------------------------------------------------------------- reference arm

def host_threads(ref_cols: int) -> int:
    """the reference allocates a 4*q*rlen byte cost matrix per worker thread (sigfish.c:873)"""
    n = os.cpu_count() or 1
    try:
        avail = int(re.search(r"MemAvailable:\s+(\d+)", open("/proc/meminfo").read()).group(1)) * 1024
        per = 4 * Q * (ref_cols // 2) * 1.25
        n = max(1, min(n, int(avail * 0.7 // per)))
    except Exception:
        pass
    return n


def run_reference_once(workdir: str, n_reads: int, threads: int):
    """one run of the unmodified reference binary on the sample; returns (cells, process_db seconds, rows)"""
    binp = os.path.join(ROOT, "oracle", "_ref", "sigfish")
    cmd = [binp, "dtw", os.path.join(workdir, "ref.fa"), os.path.join(workdir, "reads.slow5"), "--kmer-model",
           os.path.join(workdir, "model.txt"), "-t", str(threads), "-K", str(max(n_reads, 1)), "-o",
           os.path.join(workdir, "cpu.paf")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference binary failed: " + r.stderr[-1000:])
    t = float(re.search(r"Data processing time: ([0-9.]+) sec", r.stderr).group(1))
    rows = sum(1 for _ in open(os.path.join(workdir, "cpu.paf")))
    return t, rows


def prepare_reference_sample(n_reads: int):
    from sigfish_b200 import synth
    d = os.path.join(synth.tmpdir(), "bench_ref")
    os.makedirs(d, exist_ok=True)
    mean, stdv, seq, sigs = make_workload(n_reads, seed=1234)
    synth.write_model_file(os.path.join(d, "model.txt"), KMER, mean, stdv)
    synth.write_fasta(os.path.join(d, "ref.fa"), ["chrS"], [seq])
    synth.write_slow5_ascii(os.path.join(d, "reads.slow5"), [f"read{i}" for i in range(n_reads)], sigs,
                            kit="sqk-lsk114")
    return d, mean, seq, sigs


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    binp = os.path.join(ROOT, "oracle", "_ref", "sigfish")
    if not os.path.exists(binp):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/sigfish was not built (no /root/reference at build time)"}), file=OUT, flush=True)
        return
    ref_cols = 2 * (REF_LEN + 1 - KMER)
    threads = host_threads(ref_cols)
    n_reads = args.cpu_reads or max(threads, 8)
    d, _, _, _ = prepare_reference_sample(n_reads)
    cells_per_read = Q * ref_cols
    times = []
    for i in range(args.warmup + args.steps):
        t, rows = run_reference_once(d, n_reads, threads)
        if i >= args.warmup:
            times.append(t)
    sec = float(np.mean(times))
    gcups = n_reads * cells_per_read / sec / 1e9
    sample = f"{n_reads} reads of the workload per step ({n_reads * cells_per_read:.3g} cells), -t {threads}, process_db wall time"
    line = {"impl": "reference", "metric": "sDTW GCUPS", "value": gcups, "unit": "GCUPS", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "reads_per_s": n_reads / sec,
            "config": {"workload": WORKLOAD, "reads_per_step": n_reads, "query_size": Q, "prefix_size": P,
                       "ref_columns": ref_cols, "host": "unmodified reference binary, CPU only"},
            "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=OUT, flush=True)


# ---------------------------------------------------------------------------------- our arm

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=0,
                    help="reads per GPU per step (0: four full waves of DTW tasks, 4 x sfgpu_wave_reads() = 16576 on a 148-SM B200)")
    ap.add_argument("--ref-len", type=int, default=REF_LEN)
    ap.add_argument("--cpu-reads", type=int, default=0, help="reads in the CPU baseline sample (0: one per host thread)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    from sigfish_b200 import capi, ranks

    rank, world, local = ranks.env_rank()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    R = ranks.Ranks(backend="nccl", device=torch.device("cuda", local))
    barrier = R.barrier

    from sigfish_b200 import synth
    mean, stdv = synth.make_model(KMER)
    seq = synth.random_sequence(args.ref_len, np.random.default_rng(1))
    ctx = capi.Context(mean, KMER, flags=0, query_size=Q, prefix_size=P, device=local, n_slots=2,
                       warm_blocks=args.warm_blocks, piece_periods=args.piece_periods)
    ctx.set_ref([seq])
    # a task (one read x one strand of the 1 Mb contig) runs ~190 ms, so the batch is sized to whole waves of
    # resident warps; four waves per step (measured per-wave time: 1 wave 197 ms, 2: 209, 3: 196, 4: 193, 6: 191, 8: 194)
    n_reads = args.reads if args.reads > 0 else 4 * ctx.wave_reads
    sigs, _ = synth.simulate_reads([seq], KMER, mean, n_reads, seed=100 + rank, bases_per_read=450)
    sc = [synth.DNA_SCALING] * len(sigs)
    packed = ctx.pack(sigs, sc)
    ref_cols = ctx.ref_columns
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    # ---- device-resident throughput ----
    ctx.submit(0, *packed)
    first = ctx.collect(0).copy()
    for _ in range(args.warmup):
        ctx.resubmit(0)
        ctx.collect(0)
    sampler = ClockSampler(local)
    if rank == 0:  # one nvidia-smi poller per job, not per rank
        sampler.start()
    barrier()
    tot = dtw = evt = trc = 0.0
    cells = 0.0
    launches = 0
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
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
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    sampler.stop_flag = True
    if rank == 0:
        sampler.join()
    last = ctx.collect(0)
    assert last.tobytes() == first.tobytes(), "results changed between steps"
    mapped = int((last["qlen"] > 0).sum())

    ms_step = tot / args.steps
    # ---- end to end through the C-ABI with host buffers (double-buffered slots) ----
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        ctx.submit(i & 1, *packed)
        if i > 0:
            ctx.collect((i - 1) & 1)
    ctx.collect((args.steps - 1) & 1)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    h2d = int(packed[0].nbytes + (2 * len(sigs) + 1) * 8 + 3 * 4 * len(sigs))
    d2h = int(len(sigs) * (40 + 32))

    # slowest rank sets the step time; work is summed over ranks
    ms_step, e2e_ms, dtw_ms, evt_ms, trc_ms, wall_step = R.max(
        [ms_step, e2e_ms, dtw / args.steps, evt / args.steps, trc / args.steps, wall_ms / args.steps])
    job_cells, job_reads, job_samples = R.sum([cells, float(len(sigs)), float(sum(len(s) for s in sigs))])

    if rank == 0:
        clocks = sampler.result()
        peaks = measured_peaks()
        sm_count = torch.cuda.get_device_properties(local).multi_processor_count
        clk = (clocks["sm_mhz"] or peaks.get("sm_max_mhz", 1965.0)) * 1e6
        # issue-slot roofline of the DTW kernel: one warp instruction per scheduler per cycle
        peak_cells = sm_count * 4 * clk / ISSUE_SLOTS_PER_STEP * (32 * ROWS_PER_LANE)
        peak_naive = sm_count * 4 * clk / SASS_PER_STEP * (32 * ROWS_PER_LANE)
        # padded rows are issued but do no algorithmic work: count only qlen of the 32*R rows
        dtw_cells_per_s = cells / (dtw_ms * 1e-3)
        value = ranks.job_throughput(job_cells, ms_step)
        line = {
            "metric": "sDTW GCUPS", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "reads_per_s": job_reads / (ms_step * 1e-3),
            "config": {"workload": WORKLOAD, "reads_per_step_per_gpu": len(sigs), "query_size": Q, "prefix_size": P,
                       "kmer": KMER, "ref_columns": int(ref_cols), "cells_per_step_per_gpu": cells,
                       "samples_per_step_per_gpu": int(sum(len(s) for s in sigs)), "mapped_reads": mapped,
                       "batch": "four full waves of (read, strand) DTW tasks per step (4 x sfgpu_wave_reads)" if args.reads <= 0 else "--reads",
                       "l2": "256 MB buffer written between timed iterations (L2 flush)",
                       "parallelism": f"reads sharded over {world} GPU(s), reference replicated, no collective"},
            "split": split,
            "stage_ms": {"events": evt_ms, "dtw": dtw_ms, "merge_trace": trc_ms, "wall_per_step_incl_flush": wall_step},
            "e2e": {"value": job_cells / (e2e_ms * 1e-3) / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "reads_per_s": job_reads / (e2e_ms * 1e-3)},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "alu-issue", "kernel": "sf_dtw_pair_kernel<16,false,9>",
                         "achieved": dtw_cells_per_s / 1e9, "peak": peak_cells / 1e9, "unit": "GCUPS",
                         "frac": dtw_cells_per_s / peak_cells,
                         "traffic": NCU_DTW_TRAFFIC["bytes"] * len(sigs) / NCU_DTW_TRAFFIC["reads"],
                         "traffic_note": "bytes per launch, scaled by reads from the ncu capture of this workload at 8288 reads "
                                         "(profiles/r01_ncu_summary.md); "
                                         "algorithmic HBM bytes are 0.016 B/cell (8 MB reference stream, L2 resident) "
                                         "plus the wavefront checkpoints (0.59 MB/read)",
                         "peak_source": f"{sm_count} SMs x 4 schedulers x {clk / 1e6:.0f} MHz (median under load) / "
                                        f"{ISSUE_SLOTS_PER_STEP:.0f} issue slots per 32x{ROWS_PER_LANE} cells "
                                        f"({SASS_PER_STEP:.0f} SASS per column, the {ROWS_PER_LANE} half-rate FMNMX3 counted twice)",
                         "frac_of_recurrence_floor": dtw_cells_per_s / (sm_count * 4 * clk * 32 / 4.0),
                         "frac_if_every_sass_were_one_slot": dtw_cells_per_s / peak_naive,
                         "events_kernel_GBps": (job_samples / world) * 2 / (evt_ms * 1e-3) / 1e9 if evt_ms > 0 else None,
                         "hbm_peak_GBps": peaks.get("hbm_gbs")},
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                binp = os.path.join(ROOT, "oracle", "_ref", "sigfish")
                if os.path.exists(binp):
                    threads = host_threads(int(ref_cols))
                    n_cpu = args.cpu_reads or max(threads, 8)
                    d, _, _, _ = prepare_reference_sample(n_cpu)
                    sec, rows = run_reference_once(d, n_cpu, threads)
                    line["cpu_baseline"] = {"value": n_cpu * Q * ref_cols / sec / 1e9, "unit": "GCUPS", "cores": threads,
                                            "kind": "reference", "reads_per_s": n_cpu / sec,
                                            "sample": f"{n_cpu} reads of the workload, one run, -t {threads}, process_db wall {sec:.2f} s"}
                else:
                    line["cpu_baseline"] = {"value": None, "unit": "GCUPS", "cores": 0, "kind": "reference",
                                            "sample": "oracle/_ref/sigfish missing"}
            except Exception as e:  # the baseline is reported, never allowed to sink the bench line
                line["cpu_baseline"] = {"value": None, "unit": "GCUPS", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}
        print(json.dumps(line), file=OUT, flush=True)
    ctx.close()
    R.close()


if __name__ == "__main__":
    main()
