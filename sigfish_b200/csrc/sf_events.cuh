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
// One read per WARP for the parallel phases, ONE warp per block for the serial phase.  The signal is consumed in
// tiles of SF_EV_TILE samples:
//   1. every lane loads 8 consecutive int16 samples with one 16-byte load and converts to pA;
//   2. warp-wide fp64 scan of (x, x*x) continuing from the previous tile.  Every addition is
//      checked with TwoSum: when all partial sums are exact the scan equals the reference's
//      sequential sums bit for bit; a tile with any inexact addition is redone sequentially by
//      one thread in the reference's order;
//   3. every lane computes both t-statistics for its positions (32 samples behind the scan so
//      that the right-hand window is available);
//   4. lane 0 runs the two coupled peak finders over the tile (inherently sequential), closes
//      events as boundaries are emitted and stops the warp as soon as the query window is
//      complete (exact: the detector is causal).  The finder state lives in shared memory
//      between tiles, so the parallel phases do not carry it in registers (96 registers, 5 blocks per SM).
// Measured dead end (round 2, NOTES.md): running the detectors of all the block's reads on the lanes of ONE warp
// (a quarter of the instructions) is slower, 4.9 - 6.7 ms against 3.3 ms per 16 384 reads: the detector is a
// latency chain, and divergent lanes serialise it further; what hides it is many warps per SM.
#pragma once
#include <cuda_runtime.h>
#include <cfloat>
#include "sf_types.cuh"

#ifndef SF_EV_READS_PER_BLOCK
#define SF_EV_READS_PER_BLOCK 4                       // reads (= warps) per block
#endif
#ifndef SF_EV_MIN_BLOCKS
#define SF_EV_MIN_BLOCKS 5                            // 96 registers; 4 - 7 blocks measured, 5 is the fastest
#endif
#define SF_EV_THREADS (32 * SF_EV_READS_PER_BLOCK)
#define SF_EV_PER_THREAD 8
#define SF_EV_TILE (32 * SF_EV_PER_THREAD)           // 256 samples per warp per tile
#define SF_EV_LAG 32                                  // t-stat runs this far behind the scan
#define SF_EV_KEEP 64                                 // prefix sums kept from the previous tile
#define SF_EV_S_ROW (SF_EV_KEEP + SF_EV_TILE + 1)    // doubles per read: 2 x 321 words, lanes k hit banks 2k (mod 32)
#define SF_EV_T_ROW (SF_EV_TILE + SF_EV_LAG)         // floats per read

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

// Shared-memory reads of the serial loop go through 32-bit shared addresses made opaque to the compiler: it otherwise
// re-derives the buffer addresses (and the loop bound) from their parts in every iteration -- half of the loop's
// instructions -- to save registers.
__device__ __forceinline__ unsigned sf_ev_saddr(const void *p)
{
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("" : "+r"(a));
    return a;
}
template <int OFF> __device__ __forceinline__ float sf_ev_lds(unsigned a)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(a), "n"(OFF) : "memory");
    return v;
}
__device__ __forceinline__ double sf_ev_lds64(unsigned a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}

// everything the serial phase of one read carries from tile to tile
struct sf_ev_state {
    sf_finder f0, f1;
    long long npk;              // boundaries found = events closed
    long long need_peaks;       // events needed before the read may stop: boundary index qend-1 must exist
    long long qs;               // automatic start: index of the first event at or after the poly-A end (-1: not yet)
    unsigned long long prev_b;  // start of the open event
    double prev_s;              // prefix sum there
    int stop;                   // query window complete
    int pad;
};

__host__ __device__ inline size_t sf_events_smem_bytes()
{
    return (size_t)SF_EV_READS_PER_BLOCK * (2 * SF_EV_S_ROW * sizeof(double) + 2 * SF_EV_T_ROW * sizeof(float) + sizeof(sf_ev_state)) + 16;
}

__global__ void __launch_bounds__(SF_EV_THREADS, SF_EV_MIN_BLOCKS) sf_events_kernel(const sf_ev_args a)
{
    constexpr int K = SF_EV_READS_PER_BLOCK;
    // per read: prefix sums of the samples [rel, rel + KEEP + TILE] (slot j holds S[rel + j]), the two statistics of
    // the tile, the detector state
    extern __shared__ double sf_ev_smem[];
    double *S_all = sf_ev_smem;
    double *SS_all = S_all + K * SF_EV_S_ROW;
    sf_ev_state *st_all = reinterpret_cast<sf_ev_state *>(SS_all + K * SF_EV_S_ROW);
    float *T1_all = reinterpret_cast<float *>(st_all + K);
    float *T2_all = T1_all + K * SF_EV_T_ROW;

    const unsigned full = 0xffffffffu;
    const int warp = threadIdx.x >> 5;
    const int tid = threadIdx.x & 31; // lane: the warp is the unit of work of the parallel phases
    const int read = blockIdx.x * K + warp;
    const bool have = read < a.n_reads;
    double *S = S_all + warp * SF_EV_S_ROW, *SS = SS_all + warp * SF_EV_S_ROW;
    float *T1 = T1_all + warp * SF_EV_T_ROW, *T2 = T2_all + warp * SF_EV_T_ROW;
    const long long n = have ? a.sig_len[read] : 0;
    const int16_t *raw = a.signal + (have ? a.sig_off[read] : 0);
    const bool rna = (a.flags & SF_RNA) != 0;
    const bool from_end = (a.flags & SF_END) != 0;
    const int w1 = rna ? 7 : 3, w2 = rna ? 14 : 6;
    const float height = rna ? 1.0f : 0.2f;
    const float unit = have ? __fdiv_rn(a.range[read], a.digitisation[read]) : 1.0f;
    const float offs = have ? a.offset[read] : 0.0f;
    const int cap = a.ev_cap;
    uint64_t *ev_start = a.ev_start + (size_t)(have ? read : 0) * cap;
    float *ev_mean = a.ev_mean + (size_t)(have ? read : 0) * cap;
    float *ev_len = a.ev_len + (size_t)(have ? read : 0) * cap;
    const bool autop = a.p < 0 && !a.keep_all;
    const long long n_tiles = (n + SF_EV_TILE - 1) / SF_EV_TILE;
    int sticky = 0;

    if (tid == 0) {
        S[SF_EV_KEEP] = 0.0; SS[SF_EV_KEEP] = 0.0; // S[0] for the first tile (rel = -KEEP)
        sf_ev_state z;
        z.f0.threshold = rna ? 2.5f : 1.4f; z.f1.threshold = 9.0f;
        z.f0.window = w1; z.f1.window = w2;
        z.f0.masked_to = 0; z.f1.masked_to = 0;
        z.f0.peak_sum = 0.0; z.f1.peak_sum = 0.0;
        sf_finder_reset(z.f0); sf_finder_reset(z.f1);
        z.npk = 0;
        z.need_peaks = (from_end || a.keep_all) ? (1ll << 62) : (long long)(a.p < 0 ? 50 : a.p) + a.q;
        if (autop && have && a.polya_end[read] > 0) // sigfish.c:380-422
            z.need_peaks = 1ll << 62; // until the first event at or after the poly-A end is known
        z.qs = -1;
        z.prev_b = 0;
        z.prev_s = 0.0;
        z.stop = 0;
        z.pad = 0;
        st_all[warp] = z;
        if (have && n <= 0) {
            sf_readinfo ri; ri.n_events = 0; ri.qstart = ri.qend = ri.qlen = 0; ri.status = 0; ri.start_raw = ri.end_raw = 0;
            a.info[read] = ri;
        }
    }
    __syncwarp();
    if (!have || n <= 0)
        return;
    const long long pe = autop ? a.polya_end[read] : -1; // sigfish.c:380-422

    for (long long tile = 0; tile < n_tiles; tile++) {
        const long long base = tile * SF_EV_TILE;  // first sample scanned in this tile
        const long long rel = base - SF_EV_KEEP;   // sample count of slot 0
        const long long lo_pos = base - SF_EV_LAG < 0 ? 0 : base - SF_EV_LAG;
        // ---- 1+2: load, convert, scan ----
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
                sticky = 1;
            }
            __syncwarp();
        }

        // ---- 3: t-statistics for positions [base - LAG, hi_pos) ----
        const long long scanned = min(base + SF_EV_TILE, n); // prefix sums known up to S[scanned]
        const long long hi_pos = (scanned >= n) ? n : base + SF_EV_TILE - SF_EV_LAG;
        for (long long i = lo_pos + tid; i < hi_pos; i += 32) {
            float t1 = 0.0f, t2 = 0.0f;
            if (n >= 2 * w1 && i >= w1 && i <= n - w1) t1 = sf_tstat_at(S, SS, i, w1, rel);
            if (n >= 2 * w2 && i >= w2 && i <= n - w2) t2 = sf_tstat_at(S, SS, i, w2, rel);
            T1[i - lo_pos] = t1; T2[i - lo_pos] = t2;
        }
        __syncwarp();

        // ---- 4: sequential peak finders + event closing (lane 0); the state comes from and returns to shared memory.
        //      Inside a tile every position is a 32-bit offset from lo_pos (the 64-bit compares and adds of absolute
        //      positions were a third of the loop's instructions). ----
        if (tid == 0) {
            sf_ev_state z = st_all[warp];
#ifdef SF_EV_DET64
            sf_finder &f0 = z.f0, &f1 = z.f1;
            // the two statistics of the next position are fetched one iteration ahead: the loads do not depend on the
            // detector state, their latency would otherwise sit in every step of the serial chain
            float nx1 = T1[0], nx2 = T2[0];
            for (long long i = lo_pos; i < hi_pos; i++) {
                const float cur1 = nx1, cur2 = nx2;
                if (i + 1 < hi_pos) {
                    nx1 = T1[i + 1 - lo_pos];
                    nx2 = T2[i + 1 - lo_pos];
                }
#pragma unroll
                for (int d = 0; d < 2; d++) {
                    sf_finder &me = d == 0 ? f0 : f1;
                    if (me.masked_to >= (unsigned long long)i)
                        continue;
                    const float v = d == 0 ? cur1 : cur2;
                    if (me.peak_pos < 0) {
                        if (v < me.peak_val) {
                            me.peak_val = v;
                        } else if (__fsub_rn(v, me.peak_val) > height) {
                            me.peak_val = v;
                            me.peak_pos = i;
                            me.peak_sum = S[i - rel];
                        }
                        continue;
                    }
                    if (v > me.peak_val) {
                        me.peak_val = v;
                        me.peak_pos = i;
                        me.peak_sum = S[i - rel];
                    }
                    if (d == 0 && me.peak_val > me.threshold) {
                        f1.masked_to = (unsigned long long)me.peak_pos + me.window;
                        sf_finder_reset(f1);
                    }
                    if (__fsub_rn(me.peak_val, v) > height && me.peak_val > me.threshold)
                        me.valid = 1;
                    if (me.valid && ((unsigned long long)i - (unsigned long long)me.peak_pos) > me.window / 2) {
                        // boundary at peak_pos closes the open event (events.c:461-477)
                        const unsigned long long b = (unsigned long long)me.peak_pos;
                        const float len = (float)(b - z.prev_b);
                        const float mean = __fdiv_rn(__double2float_rn(__dsub_rn(me.peak_sum, z.prev_s)), len);
                        if (!autop) {
                            const long long slot = z.npk % cap;
                            ev_start[slot] = z.prev_b; ev_mean[slot] = mean; ev_len[slot] = len;
                        } else {
                            if (z.npk < a.cap_a) { ev_start[z.npk] = z.prev_b; ev_mean[z.npk] = mean; ev_len[z.npk] = len; }
                            if (z.qs >= 0 && z.npk - z.qs < cap - a.cap_a) {
                                const long long slot = a.cap_a + (z.npk - z.qs);
                                ev_start[slot] = z.prev_b; ev_mean[slot] = mean; ev_len[slot] = len;
                            }
                        }
                        z.npk++;
                        z.prev_b = b; z.prev_s = me.peak_sum;
                        // the event that opens here is the first one starting at or after the poly-A end
                        if (autop && z.qs < 0 && pe > 0 && b >= (unsigned long long)pe) {
                            z.qs = z.npk;
                            z.need_peaks = z.qs + a.q;
                        }
                        me.peak_pos = -1;
                        me.peak_val = v;
                        me.valid = 0;
                    }
                }
                if (z.npk >= z.need_peaks) {
                    z.stop = 1;
                    break;
                }
            }
#else
            constexpr int FAR = -(1 << 30); // a peak that far behind the tile keeps its absolute position in the state
            const int cnt = (int)(hi_pos - lo_pos);
            const int s_off = (int)(lo_pos - rel); // S[i - rel] = S[j + s_off]
            // per finder, relative to lo_pos: positions <= msk are masked; pk = peak position (hp: there is a peak)
            int msk[2], pk[2], valid[2];
            bool hp[2];
            float pv[2];
            double psum[2];
            const float thr[2] = {z.f0.threshold, z.f1.threshold};
            const int win[2] = {w1, w2};
#pragma unroll
            for (int d = 0; d < 2; d++) {
                const sf_finder &me = d == 0 ? z.f0 : z.f1;
                const long long dm = (long long)me.masked_to - lo_pos;
                msk[d] = dm < 0 ? -1 : (int)(dm > (1 << 20) ? (1 << 20) : dm);
                hp[d] = me.peak_pos >= 0;
                const long long dp = me.peak_pos - lo_pos;
                pk[d] = dp < FAR ? FAR : (int)dp;
                pv[d] = me.peak_val;
                valid[d] = me.valid;
                psum[d] = me.peak_sum;
            }
            int slot = (int)(z.npk % cap); // ring slot of the next event (prefix_size >= 0)
            int cnt_o = cnt;
            asm volatile("" : "+r"(cnt_o));
            // loop constants held in registers (otherwise re-selected from the RNA flag at every use)
            float height_o = height;
            int half0 = win[0] / 2, half1 = win[1] / 2;
            asm volatile("" : "+f"(height_o), "+r"(half0), "+r"(half1));
            unsigned ta = sf_ev_saddr(T1);               // T2 sits K rows behind T1
            constexpr int T2_OFF = K * SF_EV_T_ROW * (int)sizeof(float);
            const unsigned sa = sf_ev_saddr(S + s_off);  // prefix sum at position j: sa + 8 j
            // the two statistics of the next position are fetched one iteration ahead: the loads do not depend on the
            // detector state, their latency would otherwise sit in every step of the serial chain (the fetch past the
            // last position reads a slot nobody uses: the rows are followed by the next row or the pad of the buffer)
            float nx1 = sf_ev_lds<0>(ta), nx2 = sf_ev_lds<T2_OFF>(ta);
            for (int j = 0; j < cnt_o; j++) {
                const float cur1 = nx1, cur2 = nx2;
                ta += 4;
                nx1 = sf_ev_lds<0>(ta);
                nx2 = sf_ev_lds<T2_OFF>(ta);
                bool closed = false;
#pragma unroll
                for (int d = 0; d < 2; d++) {
                    if (j <= msk[d])
                        continue;
                    const float v = d == 0 ? cur1 : cur2;
                    if (!hp[d]) {
                        if (v < pv[d]) {
                            pv[d] = v;
                        } else if (__fsub_rn(v, pv[d]) > height_o) {
                            pv[d] = v;
                            hp[d] = true;
                            pk[d] = j;
                            psum[d] = sf_ev_lds64(sa + 8 * j);
                        }
                        continue;
                    }
                    if (v > pv[d]) {
                        pv[d] = v;
                        pk[d] = j;
                        psum[d] = sf_ev_lds64(sa + 8 * j);
                    }
                    if (d == 0 && pv[0] > thr[0]) { // the short finder's peak masks the long one (events.c:411-416)
                        msk[1] = pk[0] + win[0];
                        hp[1] = false;
                        pv[1] = FLT_MAX;
                        valid[1] = 0;
                    }
                    if (__fsub_rn(pv[d], v) > height_o && pv[d] > thr[d])
                        valid[d] = 1;
                    if (valid[d] && (j - pk[d]) > (d == 0 ? half0 : half1)) {
                        // boundary at the peak closes the open event (events.c:461-477)
                        const unsigned long long b = pk[d] == FAR ? (unsigned long long)(d == 0 ? z.f0.peak_pos : z.f1.peak_pos)
                                                                  : (unsigned long long)(lo_pos + pk[d]);
                        const float len = (float)(b - z.prev_b);
                        const float mean = __fdiv_rn(__double2float_rn(__dsub_rn(psum[d], z.prev_s)), len);
                        if (!autop) {
                            ev_start[slot] = z.prev_b; ev_mean[slot] = mean; ev_len[slot] = len;
                            slot = slot + 1 == cap ? 0 : slot + 1;
                        } else {
                            if (z.npk < a.cap_a) { ev_start[z.npk] = z.prev_b; ev_mean[z.npk] = mean; ev_len[z.npk] = len; }
                            if (z.qs >= 0 && z.npk - z.qs < cap - a.cap_a) {
                                const long long sl = a.cap_a + (z.npk - z.qs);
                                ev_start[sl] = z.prev_b; ev_mean[sl] = mean; ev_len[sl] = len;
                            }
                        }
                        z.npk++;
                        z.prev_b = b; z.prev_s = psum[d];
                        // the event that opens here is the first one starting at or after the poly-A end
                        if (autop && z.qs < 0 && pe > 0 && b >= (unsigned long long)pe) {
                            z.qs = z.npk;
                            z.need_peaks = z.qs + a.q;
                        }
                        hp[d] = false;
                        pv[d] = v;
                        valid[d] = 0;
                        closed = true;
                    }
                }
                // the number of events only changes when one closes
                if (closed && z.npk >= z.need_peaks) {
                    z.stop = 1;
                    break;
                }
            }
#pragma unroll
            for (int d = 0; d < 2; d++) {
                sf_finder &me = d == 0 ? z.f0 : z.f1;
                me.masked_to = msk[d] < 0 ? 0ull : (unsigned long long)(lo_pos + msk[d]);
                if (!hp[d])
                    me.peak_pos = -1;
                else if (pk[d] != FAR)
                    me.peak_pos = lo_pos + pk[d]; // FAR: unchanged since the tile began
                me.peak_val = pv[d];
                me.valid = valid[d];
                me.peak_sum = psum[d];
            }
#endif
            st_all[warp] = z;
        }
        __syncwarp();
        if (st_all[warp].stop)
            break;
        // keep the last KEEP+1 prefix sums for the next tile
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
    // the serial phase's results (lane 0 uses them)
    const long long npk = st_all[warp].npk;
    long long qs = st_all[warp].qs;
    const unsigned long long prev_b = st_all[warp].prev_b;
    const double prev_s = st_all[warp].prev_s;
    const int stop_flag = st_all[warp].stop;
    __syncwarp();

    // ---- window + z-score + query (fp32 sums in the reference's order by lane 0; everything elementwise by the warp) ----
    sf_readinfo ri;
    ri.status = 0; ri.n_events = 0; ri.qstart = ri.qend = ri.qlen = 0; ri.start_raw = ri.end_raw = 0;
    long long lo = 0, hi = 0;
    if (tid == 0) {
        ri.status = sticky ? 8 : 0;
        long long nev;
        if (stop_flag) {
            nev = npk + 1; // lower bound; only compared against qend, which it exceeds
        } else if (npk == 0) {
            nev = 0;       // the reference reads peaks[-1] here (events.c:504): undefined
            ri.status |= 4;
        } else {
            // last event runs to the end of the signal (events.c:503-505); S[n] is in the last tile
            const long long rel = (n_tiles - 1) * SF_EV_TILE - SF_EV_KEEP;
            const float len = (float)((unsigned long long)n - prev_b);
            const float mean = __fdiv_rn(__double2float_rn(__dsub_rn(S[n - rel], prev_s)), len);
            if (!autop) {
                const long long slot = npk % cap;
                ev_start[slot] = prev_b; ev_mean[slot] = mean; ev_len[slot] = len;
            } else {
                if (npk < a.cap_a) { ev_start[npk] = prev_b; ev_mean[npk] = mean; ev_len[npk] = len; }
                if (qs >= 0 && npk - qs < cap - a.cap_a) {
                    const long long slot = a.cap_a + (npk - qs);
                    ev_start[slot] = prev_b; ev_mean[slot] = mean; ev_len[slot] = len;
                }
            }
            nev = npk + 1;
        }
        ri.n_events = nev;
        long long nn = nev;
        if (nn > 0) {
            if (!from_end) { // sigfish.c:435-462
                lo = a.p;
                if (a.p < 0) {
                    if (qs >= 0) {
                        lo = qs;
                    } else { // detector failed, or no event starts after the poly-A end: 50-event fall-back
                        lo = 50;
                        ri.status |= 16;
                    }
                }
                hi = lo + a.q;
                if (lo + 25 > nn) { lo = hi = 0; nn = 0; ri.status |= 1; }
                else if (hi > nn) { hi = nn; ri.status |= 2; }
            } else {         // sigfish.c:464-478
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
    }
    // the window's events are read back by the whole warp (the stores above came from lane 0)
    __syncwarp();
    lo = __shfl_sync(full, lo, 0);
    hi = __shfl_sync(full, hi, 0);
    qs = __shfl_sync(full, qs, 0);
    const int qlen = __shfl_sync(full, ri.qlen, 0);
    if (qlen > 0) {
        // where event j of the window lives: ring slot, or in automatic mode the detected-start area
        const long long base_q = (autop && qs >= 0) ? (long long)a.cap_a - qs : 0;
        const bool ring = !autop;
#define SF_EV_SLOT(j) (ring ? (j) % cap : (j) + base_q)
        // the prefix-sum buffers are free now: the window's event means are staged there (2 x 642 floats, enough for
        // the largest query), so that the serial sums below run from shared memory
        float *wm1 = reinterpret_cast<float *>(S), *wm2 = reinterpret_cast<float *>(SS);
        const int half_cap = 2 * (SF_EV_KEEP + SF_EV_TILE + 1);
        const bool staged = qlen <= 2 * half_cap;
        if (staged) {
            for (int k = tid; k < qlen; k += 32) {
                const float m = ev_mean[SF_EV_SLOT(lo + k)];
                if (k < half_cap) wm1[k] = m; else wm2[k - half_cap] = m;
            }
        }
        __syncwarp();
        auto mean_at = [&](int k) -> float {
            if (staged)
                return k < half_cap ? wm1[k] : wm2[k - half_cap];
            return ev_mean[SF_EV_SLOT(lo + k)];
        };
        // sigfish.c:483-502: sequential fp32 sums
        float mean = 0.0f, sd = 0.0f;
        if (tid == 0) {
            const float cnt = (float)qlen;
            for (int k = 0; k < qlen; k++) mean = __fadd_rn(mean, mean_at(k));
            mean = __fdiv_rn(mean, cnt);
            float var = 0.0f;
            for (int k = 0; k < qlen; k++) {
                const float d = __fsub_rn(mean_at(k), mean);
                var = __fadd_rn(var, __fmul_rn(d, d));
            }
            var = __fdiv_rn(var, cnt);
            sd = __fsqrt_rn(var);
        }
        mean = __shfl_sync(full, mean, 0);
        sd = __shfl_sync(full, sd, 0);
        float *qv = a.queries + (size_t)read * a.q_cap;
        const bool flip = rna && !(a.flags & SF_INV); // sigfish.c:857-867
        for (int k = tid; k < qlen; k += 32) {
            const float z = __fdiv_rn(__fsub_rn(mean_at(k), mean), sd);
            qv[flip ? qlen - 1 - k : k] = z;
            if (a.win_start) {
                a.win_start[(size_t)read * a.q_cap + k] = ev_start[SF_EV_SLOT(lo + k)];
                a.win_len[(size_t)read * a.q_cap + k] = ev_len[SF_EV_SLOT(lo + k)];
            }
        }
        if (tid == 0) {
            // sigfish.c:804-805 (uint64 + float evaluates in fp32)
            ri.start_raw = ev_start[SF_EV_SLOT(lo)];
            const long long le = hi - 1;
            ri.end_raw = (uint64_t)__fadd_rn((float)ev_start[SF_EV_SLOT(le)], ev_len[SF_EV_SLOT(le)]);
        }
#undef SF_EV_SLOT
    }
    if (tid == 0)
        a.info[read] = ri;
}
