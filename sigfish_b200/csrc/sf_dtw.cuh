// sf_dtw.cuh -- kernel #3: warp-wavefront subsequence / standard DTW with the chunked
// best / second-best reduction fused in.
//
// Replaces (reference, paths relative to /root/reference):
//   src/cdtw.c:171-189  subsequence()          -- the qlen x rlen min-plus recurrence
//   src/cdtw.c:69-94    std_dtw()              -- --dtw-std variant
//   src/sigfish.c:891-901, 938-948             -- last row cut into chunks of qlen columns,
//                                                 first strict minimum of each chunk
//   src/sigfish.c:575-596 update_aln()         -- insertion list; of equal scores the later wins
//
// Mapping to the hardware: one warp owns one (read, segment group) task.  Lane l keeps query rows
// [l*R, (l+1)*R) of the current DTW column in registers (R = ceil(q/32)); at step t lane l works
// on stream column t - l, so the anti-diagonal dependency is one __shfl_up_sync per step.  The
// reference events reach the lanes through a 128-float mirrored ring in shared memory (one
// coalesced global load per 32 steps per warp, one LDS per step per lane).  No cost matrix
// exists anywhere: the state is O(qlen) registers per warp.
//
// Arithmetic per cell (bit-exact with the CPU, -fmad=false):
//     t  = x_i - y_j                (FADD)
//     m  = min(min(up, diag), left) (FMNMX3)
//     D  = |t| + m                  (FADD with |.| source modifier)
// Borders: lane 0 is fed up = diag = +0 (subsequence: virtual row -1 is 0) or +INF with a single
// 0 on the diagonal in front of each segment (standard DTW); a +INF sentinel column in front of
// every segment gives the +INF virtual column -1.
#pragma once
#include <cuda_runtime.h>
#include "sf_types.cuh"

#define SF_DTW_WARPS 4
#define SF_DTW_THREADS (SF_DTW_WARPS * 32)
#define SF_RING 128

struct sf_dtw_args {
    const float *stream;
    const sf_seg *segs;
    const sf_group *groups;
    const int32_t *order;   // group ids, longest first
    int32_t n_groups;
    int32_t n_reads;
    const float *queries;   // [n_reads][q_cap]
    const sf_readinfo *info;
    int32_t q_cap;
    sf_taskres *res;        // [n_reads][n_groups]
    float *ckpt;            // [(read * ck_per_read + ck_prefix + k)][R+1][32]
    int64_t ck_per_read;
    unsigned int *counter;
};

__host__ __device__ inline int sf_smem_floats_per_warp(int R) { return SF_RING + 32 * R; }

// register of the last query row for which the 32-step block is specialised (the generic block
// stores all R registers of the last-row lane; the specialised one stores just this one): the
// common full-length queries of the default -q values (250, 500, 100) and q = multiples of R
__host__ __device__ constexpr int sf_fast_rq(int R) { return R == 8 ? 1 : (R == 16 ? 3 : (R == 4 ? 3 : R - 1)); }

__host__ __device__ constexpr int sf_dtw_min_blocks(int R) { return R <= 8 ? 10 : (R <= 12 ? 8 : (R <= 16 ? 6 : (R <= 24 ? 4 : 3))); }

// Shared-memory accesses of the hot loop go through explicit 32-bit shared addresses whose base is made
// opaque once per block of 32 steps: otherwise the compiler re-derives the ring / buffer address from
// its parts on every step (3-4 extra integer instructions per step) to save two registers.
__device__ __forceinline__ unsigned sf_smem_addr(const void *p)
{
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("" : "+r"(a));
    return a;
}
__device__ __forceinline__ float sf_lds(unsigned addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sf_sts(unsigned addr, float v)
{
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v));
}
__device__ __forceinline__ void sf_sts4(unsigned addr, float a, float b, float c, float d)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d));
}

// 32 wavefront steps.  RQ >= 0: the last query row sits in register RQ of lane lq and only that value
// is handed to the chunk scan (last[s]); RQ < 0: all R registers are stored (last[s*R + r]).
template <int R, bool STD, int RQ>
__device__ __forceinline__ void sf_dtw_block(const float (&x)[R], float (&L)[R], float &bot, float &dprev,
                                             const float *yb, float *last, const bool is_lq, const int lane,
                                             const int nz)
{
    const unsigned full = 0xffffffffu;
    const unsigned yb_s = sf_smem_addr(yb);
    const unsigned last_s = sf_smem_addr(last);
#pragma unroll
    for (int s = 0; s < 32; s++) {
        const float yy = sf_lds(yb_s + 4 * s);
        float up = __shfl_up_sync(full, bot, 1);
        if (STD) {
            if (lane == 0)
                up = yy == SF_INF ? 0.0f : SF_INF;
        } else {
            // lane 0 is fed +0 (virtual row -1); integer multiply keeps this off the half-rate ALU pipe
            int ub;
            asm("mul.lo.s32 %0, %1, %2;" : "=r"(ub) : "r"(__float_as_int(up)), "r"(nz)); // IMAD: fma pipe
            up = __int_as_float(ub);
        }
        const float unext = up;
        float dg = dprev;
        float t[R];
#pragma unroll
        for (int r = 0; r < R; r++)
            t[r] = x[r] - yy;
#pragma unroll
        for (int r = 0; r < R; r++) {
            const float m = fminf(fminf(up, dg), L[r]);
            const float nv = fabsf(t[r]) + m;
            dg = L[r];
            L[r] = nv;
            up = nv;
        }
        dprev = unext;
        bot = L[R - 1];
        if (is_lq) {
            if (RQ >= 0) {
                sf_sts(last_s + 4 * s, L[RQ]);
            } else if (R % 4 == 0) {
#pragma unroll
                for (int r = 0; r < R; r += 4)
                    sf_sts4(last_s + 4 * (s * R + r), L[r], L[r + 1], L[r + 2], L[r + 3]);
            } else {
#pragma unroll
                for (int r = 0; r < R; r++)
                    sf_sts(last_s + 4 * (s * R + r), L[r]);
            }
        }
    }
}

template <int R, bool STD>
__global__ void __launch_bounds__(SF_DTW_THREADS, sf_dtw_min_blocks(R)) sf_dtw_score_kernel(const sf_dtw_args a)
{
    extern __shared__ float sf_smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float *ring = sf_smem + warp * sf_smem_floats_per_warp(R);
    float *last = ring + SF_RING;
    const unsigned full = 0xffffffffu;
    const unsigned n_tasks = (unsigned)a.n_groups * (unsigned)a.n_reads;

    for (;;) {
        unsigned task = 0;
        if (lane == 0)
            task = atomicAdd(a.counter, 1u);
        task = __shfl_sync(full, task, 0);
        if (task >= n_tasks)
            break;
        const int gi = task / (unsigned)a.n_reads;
        const int read = task - gi * a.n_reads;
        const int gid = a.order[gi];
        const sf_group grp = a.groups[gid];
        sf_taskres *out = a.res + (size_t)read * a.n_groups + gid;
        const int qlen = a.info[read].qlen;
        if (qlen <= 0) {
            if (lane == 0) {
                out->s1 = SF_INF; out->s2 = SF_INF; out->seg = -1; out->chunk = 0; out->pos = -1;
            }
            continue;
        }
        const int lq = (qlen - 1) / R; // lane holding the last query row
        const int rq = (qlen - 1) % R; // its register
        const bool is_lq = lane == lq;
        const bool fast = rq == sf_fast_rq(R); // warp-uniform
        const int nz = lane != 0;

        // query rows of this lane; rows past qlen are padding (finite, never read back)
        float x[R], L[R];
        const float *q = a.queries + (size_t)read * a.q_cap;
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int row = lane * R + r;
            x[r] = row < qlen ? q[row] : 0.0f;
            L[r] = SF_INF;
        }
        float bot = SF_INF;
        float dprev = (lane == 0 && !STD) ? 0.0f : SF_INF;

        const float *y = a.stream + grp.begin;
        const int n_pos = (int)(grp.end - grp.begin); // includes the leading sentinel
        const int n_blocks = (n_pos + lq + 31) >> 5;

        // ring: block b lives at slots (b&1)*32 + j and again 64 slots later
        {
            const float y0 = lane < n_pos ? y[lane] : SF_INF;
            ring[lane] = y0; ring[64 + lane] = y0;
            ring[32 + lane] = SF_INF; ring[96 + lane] = SF_INF;
        }
        __syncwarp();

        // chunk bookkeeping (warp-uniform)
        int si = grp.seg0;
        const int si_end = grp.seg0 + grp.nseg;
        sf_seg seg = a.segs[si];
        int lo = (int)(seg.off - grp.begin), hi = lo + seg.rlen;
        int chunk = 0;
        int clo = STD ? hi - 1 : lo;
        int chi = STD ? hi : min(lo + qlen, hi);
        float rmin = SF_INF; // per lane running minimum of the open chunk
        int rpos = -1;
        float s1 = SF_INF, s2 = SF_INF;
        int bseg = -1, bchunk = 0, bpos = -1;
        int ck = 0;

        for (int b = 0; b < n_blocks; b++) {
            // prefetch the next 32 reference events; consumed after the 32 steps below
            const int nidx = 32 * (b + 1) + lane;
            const float ynext = nidx < n_pos ? __ldg(y + nidx) : SF_INF;
            const float *yb = ring + ((b & 1) ? 32 : 64) - lane;

            if (fast)
                sf_dtw_block<R, STD, sf_fast_rq(R)>(x, L, bot, dprev, yb, last, is_lq, lane, nz);
            else
                sf_dtw_block<R, STD, -1>(x, L, bot, dprev, yb, last, is_lq, lane, nz);
            __syncwarp();

            // ---- last-row chunk minima (sigfish.c:891-901) ----
            {
                const int p0 = 32 * b - lq;
                const int pos = p0 + lane;
                const float v = fast ? last[lane] : last[lane * R + rq];
                for (;;) {
                    if (pos >= clo && pos < chi && v < rmin) {
                        rmin = v;
                        rpos = pos;
                    }
                    if (chi > p0 + 32)
                        break;
                    // chunk complete: first strict minimum = smallest value, then smallest column
                    float m = rmin;
                    int mp = rpos;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const float om = __shfl_xor_sync(full, m, o);
                        const int op = __shfl_xor_sync(full, mp, o);
                        if (om < m || (om == m && (unsigned)op < (unsigned)mp)) {
                            m = om;
                            mp = op;
                        }
                    }
                    // update_aln(): a later candidate with an equal score ranks better
                    if (m <= s1) {
                        s2 = s1; s1 = m; bseg = si; bchunk = chunk; bpos = mp >= 0 ? mp - lo : -1;
                    } else if (m < s2) {
                        s2 = m;
                    }
                    rmin = SF_INF;
                    rpos = -1;
                    clo = chi;
                    chunk++;
                    if (clo >= hi) {
                        si++;
                        chunk = 0;
                        if (si < si_end) {
                            seg = a.segs[si];
                            lo = (int)(seg.off - grp.begin);
                            hi = lo + seg.rlen;
                            clo = STD ? hi - 1 : lo;
                        } else {
                            clo = 0x7fffffff;
                            hi = 0x7fffffff;
                        }
                    }
                    chi = (clo == 0x7fffffff) ? 0x7fffffff : (STD ? hi : min(clo + qlen, hi));
                }
            }

            // ---- checkpoint of the skewed wavefront (for the start-coordinate pass) ----
            if (grp.ck_every > 0 && ck < grp.n_ck && (b + 1) == (ck + 1) * grp.ck_every) {
                float *c = a.ckpt + ((size_t)read * a.ck_per_read + grp.ck_prefix + ck) * (size_t)((R + 1) * 32);
#pragma unroll
                for (int r = 0; r < R; r++)
                    c[r * 32 + lane] = L[r];
                c[R * 32 + lane] = dprev;
                ck++;
            }

            // publish block b+1 of the reference events (overwrites block b-1)
            {
                const int slot = ((b + 1) & 1) * 32 + lane;
                ring[slot] = ynext;
                ring[slot + 64] = ynext;
            }
            __syncwarp();
        }

        if (lane == 0) {
            out->s1 = s1; out->s2 = s2; out->seg = bseg; out->chunk = bchunk; out->pos = bpos;
        }
    }
}
