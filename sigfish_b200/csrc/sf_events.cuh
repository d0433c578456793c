// sf_events.cuh -- kernel #1: raw int16 signal -> events -> query window -> z-scored query.
//
// Replaces (reference, paths relative to /root/reference):
//   src/sigfish.c:334-347   pA conversion (event_single)
//   src/events.c:297-307    compute_sum_sumsq   (fp64 prefix sums, fp32 square)
//   src/events.c:319-368    compute_tstat       (windows 3/6 DNA, 7/14 RNA)
//   src/events.c:375-447    short_long_peak_detector
//   src/events.c:461-508    create_event(s)     (start / length / mean)
//   src/sigfish.c:424-505   normalise_single    (window [qstart,qend) + z-score)
//   src/sigfish.c:857-867   query construction  (reversed for RNA unless --invert)
// The MAD trim of events.c:99-269 has no effect on the reference's results (its return value is
// dropped at events.c:567) and is not implemented.
//
// Two schedules over the same device functions:
//  (A) split (the default for batches mapped from the read start): sf_tstat_kernel -- one warp per read,
//      throughput work only: pA conversion, exact fp64 prefix sums and both t-statistics for the first N0
//      samples, written to HBM -- then sf_detect_kernel -- one LANE per read runs the sequential peak
//      detector over those arrays, closes events, and builds the z-scored query.  The detector is the critical
//      path of a read, so 32 of them per warp keep every read of the batch in flight at once.  Reads that need
//      more than N0 samples are flagged and redone by (B).
//  (B) fused: sf_events_kernel -- one warp per read, tiles of SF_EV_TILE samples, lane 0 runs the detector
//      in shared memory and the warp stops as soon as the query window is complete.  Used for --from-end,
//      for event-table dumps and as the fall-back of (A).
// Per tile (both schedules):
//   1. every lane loads 8 consecutive int16 samples with one 16-byte load and converts to pA;
//   2. warp-wide fp64 scan of (x, x*x) continuing from the previous tile.  Every addition is
//      checked with TwoSum: when all partial sums are exact the scan equals the reference's
//      sequential sums bit for bit; a tile with any inexact addition is redone sequentially by
//      one thread in the reference's order;
//   3. every lane computes both t-statistics for its positions (32 samples behind the scan so
//      that the right-hand window is available);
//   4. lane 0 runs the two coupled peak finders over the tile (inherently sequential), closes
//      events as boundaries are emitted and stops the warp as soon as the query window is
//      complete (exact: the detector is causal).
#pragma once
#include <cuda_runtime.h>
#include <cfloat>
#include "sf_types.cuh"

#define SF_EV_READS_PER_BLOCK 4
#define SF_EV_THREADS (32 * SF_EV_READS_PER_BLOCK)
#define SF_EV_PER_THREAD 8
#define SF_EV_TILE (32 * SF_EV_PER_THREAD)           // 256 samples per warp per tile
#define SF_EV_LAG 32                                  // t-stat runs this far behind the scan
#define SF_EV_KEEP 64                                 // prefix sums kept from the previous tile

struct sf_ev_args {
    const int16_t *signal;      // all reads of the batch, concatenated (each read 16-byte aligned)
    const int64_t *sig_off;     // [n_reads + 1] offsets in samples (multiples of 8)
    const int64_t *sig_len;     // [n_reads]
    const float *digitisation, *offset, *range; // [n_reads] (already narrowed to fp32, sigfish.c:335-337)
    int32_t n_reads;
    uint32_t flags;
    int32_t q, p;
    int32_t ev_cap;             // event slots per read in the scratch below (ring when --from-end)
    uint64_t *ev_start;         // [n_reads][ev_cap]
    float *ev_mean;             // [n_reads][ev_cap]
    float *ev_len;              // [n_reads][ev_cap]
    float *queries;             // [n_reads][q_cap]
    int32_t q_cap;
    sf_readinfo *info;          // [n_reads]
    int32_t keep_all;           // 1: never exit early (event-table dumps for tests)
    // automatic query start (p < 0): raw-sample index where the poly-A tail ends (sf_qstart.cuh), -1 when
    // the detector found nothing.  Event slots are then split: [0, cap_a) holds events 0..cap_a-1 (the
    // 50-event fall-back window), [cap_a, ev_cap) the events from the detected start on.
    const int64_t *polya_end;
    int32_t cap_a;
    // --sam: start sample and length of every event of the query window, in event order (what
    // r2qevent_map_to_ss() reads from the event table, sigfish.c:737-742); null otherwise
    uint64_t *win_start;        // [n_reads][q_cap]
    float *win_len;             // [n_reads][q_cap]
    // split schedule: t-statistics and prefix sums of the first n0 samples of every read
    int32_t n0;                 // multiple of SF_EV_TILE
    float2 *tt;                 // [n_reads][n0]   (t-stat short window, long window)
    double *ss;                 // [n_reads][n0+1] prefix sums S[k] = sum of the first k pA samples
    int32_t *rflags;            // [n_reads] bit0: a tile needed the sequential prefix-sum redo; bit1: needs (B)
    int32_t only_flagged;       // fused kernel: process only reads with flags bit1
};

struct sf_finder {
    float threshold;
    unsigned long long window;
    unsigned long long masked_to;
    long long peak_pos; // -1: none
    float peak_val;
    int valid;
    double peak_sum; // prefix sum at peak_pos (the event boundary if this peak fires)
};

__device__ __forceinline__ void sf_finder_reset(sf_finder &f)
{
    f.peak_pos = -1;
    f.peak_val = FLT_MAX;
    f.valid = 0;
}

// events.c:343-364 for one position, explicit rounding order
__device__ __forceinline__ float sf_tstat_at(const double *S, const double *SS, long long i, int w, long long rel)
{
    // S/SS are indexed relative: S[k - rel] is the prefix sum of k samples
    const float wf = (float)w;
    double lsum = S[i - rel], lsq = SS[i - rel];
    if (i > w) {
        lsum = __dsub_rn(lsum, S[i - w - rel]);
        lsq = __dsub_rn(lsq, SS[i - w - rel]);
    }
    const float rsum = __double2float_rn(__dsub_rn(S[i + w - rel], S[i - rel]));
    const float rsq = __double2float_rn(__dsub_rn(SS[i + w - rel], SS[i - rel]));
    const float lmean = __double2float_rn(__ddiv_rn(lsum, (double)wf));
    const float rmean = __fdiv_rn(rsum, wf);
    const float lmean2 = __fmul_rn(lmean, lmean);
    const float rmean2 = __fmul_rn(rmean, rmean);
    const float rq = __fdiv_rn(rsq, wf);
    double acc = __ddiv_rn(lsq, (double)wf);
    acc = __dsub_rn(acc, (double)lmean2);
    acc = __dadd_rn(acc, (double)rq);
    acc = __dsub_rn(acc, (double)rmean2);
    float var = __double2float_rn(acc);
    var = fmaxf(var, FLT_MIN);
    const float dm = __fsub_rn(rmean, lmean);
    const float vq = __fdiv_rn(var, wf);
    return __double2float_rn(__ddiv_rn(fabs((double)dm), __dsqrt_rn((double)vq)));
}

__device__ __forceinline__ void sf_twosum(double a, double b, double &s, int &inexact)
{
    s = __dadd_rn(a, b);
    const double bb = __dsub_rn(s, a);
    const double err = __dadd_rn(__dsub_rn(a, __dsub_rn(s, bb)), __dsub_rn(b, bb));
    inexact |= (err != 0.0);
}


// ---- per-read constants and the sequential detector state ------------------------------------------------

struct sf_ev_read {
    int read;
    long long n;
    bool rna, from_end, autop;
    long long pe;       // poly-A end sample (automatic query start), -1: none
    float height;
    int cap;
    uint64_t *ev_start;
    float *ev_mean, *ev_len;
};

struct sf_det_state {
    sf_finder f0, f1;
    long long npk;            // boundaries emitted = events closed
    unsigned long long prev_b; // start of the open event
    double prev_s;            // prefix sum there
    long long qs;             // automatic start: index of the first event at or after the poly-A end
    long long need_peaks;     // events that must be closed before the read may stop
};

__device__ __forceinline__ sf_ev_read sf_ev_read_init(const sf_ev_args &a, int read)
{
    sf_ev_read r;
    r.read = read;
    r.n = a.sig_len[read];
    r.rna = (a.flags & SF_RNA) != 0;
    r.from_end = (a.flags & SF_END) != 0;
    r.autop = a.p < 0 && !a.keep_all;
    r.pe = r.autop ? a.polya_end[read] : -1; // sigfish.c:380-422
    r.height = r.rna ? 1.0f : 0.2f;
    r.cap = a.ev_cap;
    r.ev_start = a.ev_start + (size_t)read * a.ev_cap;
    r.ev_mean = a.ev_mean + (size_t)read * a.ev_cap;
    r.ev_len = a.ev_len + (size_t)read * a.ev_cap;
    return r;
}

__device__ __forceinline__ void sf_det_init(sf_det_state &d, const sf_ev_args &a, const sf_ev_read &r)
{
    d.f0.threshold = r.rna ? 2.5f : 1.4f; d.f1.threshold = 9.0f;
    d.f0.window = r.rna ? 7 : 3; d.f1.window = r.rna ? 14 : 6;
    d.f0.masked_to = 0; d.f1.masked_to = 0;
    d.f0.peak_sum = 0.0; d.f1.peak_sum = 0.0;
    sf_finder_reset(d.f0); sf_finder_reset(d.f1);
    d.npk = 0;
    d.prev_b = 0;
    d.prev_s = 0.0;
    d.qs = -1;
    // events needed before the read may stop: boundary index qend-1 must exist
    d.need_peaks = (r.from_end || a.keep_all) ? (1ll << 62) : (long long)(a.p < 0 ? 50 : a.p) + a.q;
    if (r.autop && r.pe > 0)
        d.need_peaks = 1ll << 62; // until the first event at or after the poly-A end is known
}

// event number d.npk = [start, start + len) with the given mean
__device__ __forceinline__ void sf_store_event(const sf_ev_args &a, const sf_ev_read &r, const sf_det_state &d,
                                               unsigned long long start, float mean, float len)
{
    if (!r.autop) {
        const long long slot = d.npk % r.cap;
        r.ev_start[slot] = start; r.ev_mean[slot] = mean; r.ev_len[slot] = len;
    } else {
        if (d.npk < a.cap_a) { r.ev_start[d.npk] = start; r.ev_mean[d.npk] = mean; r.ev_len[d.npk] = len; }
        if (d.qs >= 0 && d.npk - d.qs < r.cap - a.cap_a) {
            const long long slot = a.cap_a + (d.npk - d.qs);
            r.ev_start[slot] = start; r.ev_mean[slot] = mean; r.ev_len[slot] = len;
        }
    }
}

// One sample of the two coupled peak finders (events.c:375-447) + event closing (events.c:461-477).
// S_at(k) returns the prefix sum of the first k samples.  Returns true once enough events are closed.
template <typename SFn>
__device__ __forceinline__ bool sf_det_sample(const sf_ev_args &a, const sf_ev_read &r, sf_det_state &d, long long i,
                                              float v1, float v2, SFn S_at)
{
#pragma unroll
    for (int k = 0; k < 2; k++) {
        sf_finder &me = k == 0 ? d.f0 : d.f1;
        if (me.masked_to >= (unsigned long long)i)
            continue;
        const float v = k == 0 ? v1 : v2;
        if (me.peak_pos < 0) {
            if (v < me.peak_val) {
                me.peak_val = v;
            } else if (__fsub_rn(v, me.peak_val) > r.height) {
                me.peak_val = v;
                me.peak_pos = i;
                me.peak_sum = S_at(i);
            }
            continue;
        }
        if (v > me.peak_val) {
            me.peak_val = v;
            me.peak_pos = i;
            me.peak_sum = S_at(i);
        }
        if (k == 0 && me.peak_val > me.threshold) {
            d.f1.masked_to = (unsigned long long)me.peak_pos + me.window;
            sf_finder_reset(d.f1);
        }
        if (__fsub_rn(me.peak_val, v) > r.height && me.peak_val > me.threshold)
            me.valid = 1;
        if (me.valid && ((unsigned long long)i - (unsigned long long)me.peak_pos) > me.window / 2) {
            // boundary at peak_pos closes the open event
            const unsigned long long b = (unsigned long long)me.peak_pos;
            const float len = (float)(b - d.prev_b);
            const float mean = __fdiv_rn(__double2float_rn(__dsub_rn(me.peak_sum, d.prev_s)), len);
            sf_store_event(a, r, d, d.prev_b, mean, len);
            d.npk++;
            d.prev_b = b; d.prev_s = me.peak_sum;
            // the event that opens here is the first one starting at or after the poly-A end
            if (r.autop && d.qs < 0 && r.pe > 0 && b >= (unsigned long long)r.pe) {
                d.qs = d.npk;
                d.need_peaks = d.qs + a.q;
            }
            me.peak_pos = -1;
            me.peak_val = v;
            me.valid = 0;
        }
    }
    return d.npk >= d.need_peaks;
}

// Window selection + z-score + query (normalise_single, sigfish.c:424-505; query construction 857-867),
// after the detector stopped early (`stopped`) or consumed the whole read (then the last event is closed at
// n with S_n = prefix sum of all samples).  One thread; fp32 sums in the reference's order.
__device__ __forceinline__ void sf_finish_read(const sf_ev_args &a, const sf_ev_read &r, sf_det_state &d, bool stopped,
                                               int sticky, double S_n)
{
    const int cap = r.cap;
    sf_readinfo ri;
    ri.status = sticky ? 8 : 0;
    long long nev;
    if (stopped) {
        nev = d.npk + 1; // lower bound; only compared against qend, which it exceeds
    } else if (d.npk == 0) {
        nev = 0;         // the reference reads peaks[-1] here (events.c:504): undefined
        ri.status |= 4;
    } else {
        // last event runs to the end of the signal (events.c:503-505)
        const float len = (float)((unsigned long long)r.n - d.prev_b);
        const float mean = __fdiv_rn(__double2float_rn(__dsub_rn(S_n, d.prev_s)), len);
        sf_store_event(a, r, d, d.prev_b, mean, len);
        nev = d.npk + 1;
    }
    ri.n_events = nev;
    long long lo = 0, hi = 0, nn = nev;
    if (nn > 0) {
        if (!r.from_end) { // sigfish.c:435-462
            lo = a.p;
            if (a.p < 0) {
                if (d.qs >= 0) {
                    lo = d.qs;
                } else { // detector failed, or no event starts after the poly-A end: 50-event fall-back
                    lo = 50;
                    ri.status |= 16;
                }
            }
            hi = lo + a.q;
            if (lo + 25 > nn) { lo = hi = 0; nn = 0; ri.status |= 1; }
            else if (hi > nn) { hi = nn; ri.status |= 2; }
        } else {           // sigfish.c:464-478
            lo = nn - a.p - a.q;
            hi = nn - a.p;
            if (lo < 0) { lo = 0; ri.status |= 2; }
            if (hi < 0) { hi = 0; nn = 0; ri.status |= 1; }
        }
    }
    ri.qstart = (int)lo; ri.qend = (int)hi;
    int qlen = nn > 0 ? (int)(hi - lo) : 0;
    if (qlen > a.q_cap) qlen = 0; // cannot happen (hi - lo <= q)
    ri.qlen = qlen;
    ri.start_raw = 0; ri.end_raw = 0;
    if (qlen > 0) {
        // where event j of the window lives: ring slot, or in automatic mode the detected-start area
        const long long base_q = (r.autop && d.qs >= 0) ? (long long)a.cap_a - d.qs : 0;
        const bool ring = !r.autop;
#define SF_EV_SLOT(j) (ring ? (j) % cap : (j) + base_q)
        // sigfish.c:483-502
        const float cnt = (float)qlen;
        float mean = 0.0f;
        for (long long j = lo; j < hi; j++) mean = __fadd_rn(mean, r.ev_mean[SF_EV_SLOT(j)]);
        mean = __fdiv_rn(mean, cnt);
        float var = 0.0f;
        for (long long j = lo; j < hi; j++) {
            const float dv = __fsub_rn(r.ev_mean[SF_EV_SLOT(j)], mean);
            var = __fadd_rn(var, __fmul_rn(dv, dv));
        }
        var = __fdiv_rn(var, cnt);
        const float sd = __fsqrt_rn(var);
        float *qv = a.queries + (size_t)r.read * a.q_cap;
        const bool flip = r.rna && !(a.flags & SF_INV); // sigfish.c:857-867
        for (long long j = lo; j < hi; j++) {
            const float z = __fdiv_rn(__fsub_rn(r.ev_mean[SF_EV_SLOT(j)], mean), sd);
            const int k = (int)(j - lo);
            qv[flip ? qlen - 1 - k : k] = z;
            if (a.win_start) {
                a.win_start[(size_t)r.read * a.q_cap + k] = r.ev_start[SF_EV_SLOT(j)];
                a.win_len[(size_t)r.read * a.q_cap + k] = r.ev_len[SF_EV_SLOT(j)];
            }
        }
        // sigfish.c:804-805 (uint64 + float evaluates in fp32)
        ri.start_raw = r.ev_start[SF_EV_SLOT(lo)];
        const long long le = hi - 1;
        ri.end_raw = (uint64_t)__fadd_rn((float)r.ev_start[SF_EV_SLOT(le)], r.ev_len[SF_EV_SLOT(le)]);
#undef SF_EV_SLOT
    }
    a.info[r.read] = ri;
}

// One tile of the throughput part for one warp: load + pA conversion, exact fp64 prefix sums continuing from
// S[KEEP] (sequential redo of the tile when any addition was inexact), both t-statistics for the positions
// [lo_pos, hi_pos).  S/SS slot j holds the prefix sum of rel + j samples, rel = base - KEEP.
// Returns 1 when the tile needed the sequential redo.
__device__ __forceinline__ int sf_tile(const int16_t *raw, long long n, long long base, float offs, float unit, int w1,
                                       int w2, double *S, double *SS, float *T1, float *T2, int tid,
                                       long long &lo_pos, long long &hi_pos)
{
    const unsigned full = 0xffffffffu;
    const long long rel = base - SF_EV_KEEP;
    int redone = 0;
    float xs[SF_EV_PER_THREAD];
    {
        const long long i0 = base + (long long)tid * SF_EV_PER_THREAD;
        union { uint4 pk; int16_t v[SF_EV_PER_THREAD]; } u;
        if (i0 + SF_EV_PER_THREAD <= n) {
            u.pk = __ldg(reinterpret_cast<const uint4 *>(raw + i0)); // 8 samples, one 16-byte load
        } else {
#pragma unroll
            for (int k = 0; k < SF_EV_PER_THREAD; k++)
                u.v[k] = (i0 + k < n) ? raw[i0 + k] : (int16_t)0;
        }
#pragma unroll
        for (int k = 0; k < SF_EV_PER_THREAD; k++)
            xs[k] = (i0 + k < n) ? __fmul_rn(__fadd_rn((float)u.v[k], offs), unit) : 0.0f;
    }
    int inexact = 0;
    double ps[SF_EV_PER_THREAD], pq[SF_EV_PER_THREAD];
    {
        double acc = 0.0, acq = 0.0;
#pragma unroll
        for (int k = 0; k < SF_EV_PER_THREAD; k++) {
            sf_twosum(acc, (double)xs[k], acc, inexact);
            sf_twosum(acq, (double)__fmul_rn(xs[k], xs[k]), acq, inexact);
            ps[k] = acc; pq[k] = acq;
        }
    }
    // warp-inclusive scan of the lane totals
    double tot = ps[SF_EV_PER_THREAD - 1], toq = pq[SF_EV_PER_THREAD - 1];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double us = __shfl_up_sync(full, tot, o);
        const double uq = __shfl_up_sync(full, toq, o);
        if (tid >= o) {
            sf_twosum(tot, us, tot, inexact);
            sf_twosum(toq, uq, toq, inexact);
        }
    }
    double carry = S[SF_EV_KEEP], carq = SS[SF_EV_KEEP]; // prefix sum at `base`
    // exclusive prefix of this lane = carry + total of the lanes below
    double exs = __shfl_up_sync(full, tot, 1), exq = __shfl_up_sync(full, toq, 1);
    if (tid == 0) { exs = 0.0; exq = 0.0; }
    sf_twosum(carry, exs, carry, inexact);
    sf_twosum(carq, exq, carq, inexact);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < SF_EV_PER_THREAD; k++) {
        double vs, vq;
        sf_twosum(carry, ps[k], vs, inexact);
        sf_twosum(carq, pq[k], vq, inexact);
        const int slot = SF_EV_KEEP + 1 + tid * SF_EV_PER_THREAD + k;
        S[slot] = vs; SS[slot] = vq;
    }
    const int any_inexact = __any_sync(full, inexact);
    __syncwarp();
    if (any_inexact) {
        // redo this tile in the reference's order (events.c:303-306); needs the pA values again
        if (tid == 0) {
            double acc = S[SF_EV_KEEP], acq = SS[SF_EV_KEEP];
            const long long lim = min((long long)SF_EV_TILE, n - base);
            for (long long k = 0; k < lim; k++) {
                const float x = __fmul_rn(__fadd_rn((float)raw[base + k], offs), unit);
                acc = __dadd_rn(acc, (double)x);
                acq = __dadd_rn(acq, (double)__fmul_rn(x, x));
                S[SF_EV_KEEP + 1 + k] = acc; SS[SF_EV_KEEP + 1 + k] = acq;
            }
        }
        redone = 1;
        __syncwarp();
    }
    // t-statistics for positions [base - LAG, hi_pos): they trail the scan so that the right window exists
    const long long scanned = min(base + SF_EV_TILE, n); // prefix sums known up to S[scanned]
    lo_pos = base - SF_EV_LAG < 0 ? 0 : base - SF_EV_LAG;
    hi_pos = (scanned >= n) ? n : base + SF_EV_TILE - SF_EV_LAG;
    for (long long i = lo_pos + tid; i < hi_pos; i += 32) {
        float t1 = 0.0f, t2 = 0.0f;
        if (n >= 2 * w1 && i >= w1 && i <= n - w1) t1 = sf_tstat_at(S, SS, i, w1, rel);
        if (n >= 2 * w2 && i >= w2 && i <= n - w2) t2 = sf_tstat_at(S, SS, i, w2, rel);
        T1[i - lo_pos] = t1; T2[i - lo_pos] = t2;
    }
    __syncwarp();
    return redone;
}

// keep the last KEEP+1 prefix sums for the next tile
__device__ __forceinline__ void sf_tile_carry(double *S, double *SS, int tid)
{
    double ks[3], kq[3];
#pragma unroll
    for (int e = 0; e < 3; e++) {
        const int idx = tid + 32 * e;
        ks[e] = idx <= SF_EV_KEEP ? S[SF_EV_TILE + idx] : 0.0;
        kq[e] = idx <= SF_EV_KEEP ? SS[SF_EV_TILE + idx] : 0.0;
    }
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 3; e++) {
        const int idx = tid + 32 * e;
        if (idx <= SF_EV_KEEP) { S[idx] = ks[e]; SS[idx] = kq[e]; }
    }
    __syncwarp();
}

// samples whose t-statistics the split schedule provides for a read of n samples
__host__ __device__ inline long long sf_n_avail(long long n, int n0) { return n <= n0 ? n : (long long)n0 - SF_EV_LAG; }

// ---- schedule (B): fused, one warp per read ---------------------------------------------------------------

__global__ void __launch_bounds__(SF_EV_THREADS) sf_events_kernel(const sf_ev_args a)
{
    // per warp: prefix sums of the samples [rel, rel + KEEP + TILE]; slot j holds S[rel + j]
    __shared__ double S_all[SF_EV_READS_PER_BLOCK][SF_EV_KEEP + SF_EV_TILE + 1];
    __shared__ double SS_all[SF_EV_READS_PER_BLOCK][SF_EV_KEEP + SF_EV_TILE + 1];
    __shared__ float T1_all[SF_EV_READS_PER_BLOCK][SF_EV_TILE + SF_EV_LAG];
    __shared__ float T2_all[SF_EV_READS_PER_BLOCK][SF_EV_TILE + SF_EV_LAG];

    const unsigned full = 0xffffffffu;
    const int warp = threadIdx.x >> 5;
    const int tid = threadIdx.x & 31; // lane: the warp is the unit of work
    const int read = blockIdx.x * SF_EV_READS_PER_BLOCK + warp;
    if (read >= a.n_reads)
        return;
    if (a.only_flagged && !(a.rflags[read] & 2))
        return;
    double *S = S_all[warp], *SS = SS_all[warp];
    float *T1 = T1_all[warp], *T2 = T2_all[warp];
    const sf_ev_read r = sf_ev_read_init(a, read);
    const long long n = r.n;
    const int16_t *raw = a.signal + a.sig_off[read];
    const int w1 = r.rna ? 7 : 3, w2 = r.rna ? 14 : 6;
    const float unit = __fdiv_rn(a.range[read], a.digitisation[read]);
    const float offs = a.offset[read];
    sf_det_state d;
    sf_det_init(d, a, r);
    int sticky = 0;
    int stop_flag = 0;

    if (tid == 0) {
        S[SF_EV_KEEP] = 0.0; SS[SF_EV_KEEP] = 0.0; // S[0] for the first tile (rel = -KEEP)
    }
    __syncwarp();

    if (n <= 0) {
        if (tid == 0) {
            sf_readinfo ri; ri.n_events = 0; ri.qstart = ri.qend = ri.qlen = 0; ri.status = 0; ri.start_raw = ri.end_raw = 0;
            a.info[read] = ri;
        }
        return;
    }

    const long long n_tiles = (n + SF_EV_TILE - 1) / SF_EV_TILE;
    for (long long tile = 0; tile < n_tiles; tile++) {
        const long long base = tile * SF_EV_TILE;
        const long long rel = base - SF_EV_KEEP;
        long long lo_pos, hi_pos;
        sticky |= sf_tile(raw, n, base, offs, unit, w1, w2, S, SS, T1, T2, tid, lo_pos, hi_pos);
        // sequential peak finders + event closing (lane 0)
        if (tid == 0) {
            for (long long i = lo_pos; i < hi_pos; i++) {
                if (sf_det_sample(a, r, d, i, T1[i - lo_pos], T2[i - lo_pos], [&](long long k) { return S[k - rel]; })) {
                    stop_flag = 1;
                    break;
                }
            }
        }
        stop_flag = __shfl_sync(full, stop_flag, 0);
        if (stop_flag)
            break;
        sf_tile_carry(S, SS, tid);
    }
    if (tid == 0) {
        // S[n] is in the last tile when the loop ran to the end
        const long long rel = (n_tiles - 1) * SF_EV_TILE - SF_EV_KEEP;
        sf_finish_read(a, r, d, stop_flag != 0, sticky, stop_flag ? 0.0 : S[n - rel]);
    }
}

// ---- schedule (A): throughput kernel + lane-per-read detector ---------------------------------------------

__global__ void __launch_bounds__(SF_EV_THREADS) sf_tstat_kernel(const sf_ev_args a)
{
    __shared__ double S_all[SF_EV_READS_PER_BLOCK][SF_EV_KEEP + SF_EV_TILE + 1];
    __shared__ double SS_all[SF_EV_READS_PER_BLOCK][SF_EV_KEEP + SF_EV_TILE + 1];
    __shared__ float T1_all[SF_EV_READS_PER_BLOCK][SF_EV_TILE + SF_EV_LAG];
    __shared__ float T2_all[SF_EV_READS_PER_BLOCK][SF_EV_TILE + SF_EV_LAG];
    const int warp = threadIdx.x >> 5;
    const int tid = threadIdx.x & 31;
    const int read = blockIdx.x * SF_EV_READS_PER_BLOCK + warp;
    if (read >= a.n_reads)
        return;
    double *S = S_all[warp], *SS = SS_all[warp];
    float *T1 = T1_all[warp], *T2 = T2_all[warp];
    const long long n = a.sig_len[read];
    const int16_t *raw = a.signal + a.sig_off[read];
    const bool rna = (a.flags & SF_RNA) != 0;
    const int w1 = rna ? 7 : 3, w2 = rna ? 14 : 6;
    const float unit = __fdiv_rn(a.range[read], a.digitisation[read]);
    const float offs = a.offset[read];
    float2 *tt = a.tt + (size_t)read * a.n0;
    double *ss = a.ss + (size_t)read * (a.n0 + 1);
    int sticky = 0;
    if (tid == 0) {
        S[SF_EV_KEEP] = 0.0; SS[SF_EV_KEEP] = 0.0;
        ss[0] = 0.0;
    }
    __syncwarp();
    const long long n_need = n < a.n0 ? n : a.n0;
    const long long n_tiles = (n_need + SF_EV_TILE - 1) / SF_EV_TILE;
    for (long long tile = 0; tile < n_tiles; tile++) {
        const long long base = tile * SF_EV_TILE;
        long long lo_pos, hi_pos;
        sticky |= sf_tile(raw, n, base, offs, unit, w1, w2, S, SS, T1, T2, tid, lo_pos, hi_pos);
        for (long long i = lo_pos + tid; i < hi_pos; i += 32)
            tt[i] = make_float2(T1[i - lo_pos], T2[i - lo_pos]);
        const long long top = min(base + SF_EV_TILE, n); // S known up to here
        for (long long k = base + 1 + tid; k <= top; k += 32)
            ss[k] = S[k - (base - SF_EV_KEEP)];
        sf_tile_carry(S, SS, tid);
    }
    if (tid == 0)
        a.rflags[read] = sticky ? 1 : 0;
}

__global__ void __launch_bounds__(32) sf_detect_kernel(const sf_ev_args a)
{
    const int read = blockIdx.x * blockDim.x + threadIdx.x;
    if (read >= a.n_reads)
        return;
    const sf_ev_read r = sf_ev_read_init(a, read);
    if (r.n <= 0) {
        sf_readinfo ri; ri.n_events = 0; ri.qstart = ri.qend = ri.qlen = 0; ri.status = 0; ri.start_raw = ri.end_raw = 0;
        a.info[read] = ri;
        return;
    }
    const float2 *tt = a.tt + (size_t)read * a.n0;
    const double *ss = a.ss + (size_t)read * (a.n0 + 1);
    sf_det_state d;
    sf_det_init(d, a, r);
    const long long n_avail = sf_n_avail(r.n, a.n0);
    bool stopped = false;
    // the t-statistics are fetched a few samples ahead of the state machine
    constexpr int AHEAD = 4;
    float2 buf[AHEAD];
#pragma unroll
    for (int k = 0; k < AHEAD; k++)
        buf[k] = k < n_avail ? __ldg(tt + k) : make_float2(0.0f, 0.0f);
    for (long long i0 = 0; i0 < n_avail && !stopped; i0 += AHEAD) {
        float2 cur[AHEAD];
#pragma unroll
        for (int k = 0; k < AHEAD; k++) {
            cur[k] = buf[k];
            const long long nx = i0 + AHEAD + k;
            buf[k] = nx < n_avail ? __ldg(tt + nx) : make_float2(0.0f, 0.0f);
        }
#pragma unroll
        for (int k = 0; k < AHEAD; k++) {
            const long long i = i0 + k;
            if (i < n_avail && !stopped)
                stopped = sf_det_sample(a, r, d, i, cur[k].x, cur[k].y, [&](long long p) { return __ldg(ss + p); });
        }
    }
    const int sticky = a.rflags[read] & 1;
    if (!stopped && n_avail < r.n) {
        a.rflags[read] = sticky | 2; // the first n0 samples were not enough: redo with the fused kernel
        return;
    }
    sf_finish_read(a, r, d, stopped, sticky, stopped ? 0.0 : __ldg(ss + r.n));
}

