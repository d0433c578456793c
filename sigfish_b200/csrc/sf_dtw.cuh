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
// [l*R, (l+1)*R) of the DTW column in registers (R = ceil(q/32)).  A *macro-step* advances every lane
// by TWO reference columns: at macro-step T lane l computes columns 2(T-l) and 2(T-l)+1 of its rows.
// The second column of a row only needs the first column of the same row, so the two columns form two
// interleaved dependency chains (critical path R+1 cells for 2R cells: ILP 2 inside one warp), and
// the per-column overheads are shared: one 8-byte LDS for both reference events, one 8-byte STS of the
// last query row, two shuffles for the two bottom-row hand-offs to the next lane.  The reference
// events reach the lanes through a ring of float2 pairs in shared memory (one coalesced global load
// per 32 macro-steps).  No cost matrix exists anywhere: the state is O(qlen) registers per warp.
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
#define SF_RING_PAIRS 128 // float2 slots: two blocks of 32 pairs, stored twice (mirror) for wrap-free reads
#define SF_BLOCK_COLS 64  // reference columns per block of 32 macro-steps

struct sf_dtw_args {
    const float *stream;
    const sf_seg *segs;
    const sf_group *groups;
    const sf_piece *pieces; // the tasks of one read: whole groups and pieces of long segments (one split level)
    const int32_t *order;   // piece ids, longest first
    int32_t n_pieces;
    int32_t n_reads;
    const float *queries;   // [n_reads][q_cap]
    const sf_readinfo *info;
    int32_t q_cap;
    sf_taskres *res;        // [n_reads][n_pieces]
    float *ckpt;            // [(read * ck_per_read + ck_prefix + k)][R+2][32]
    int64_t ck_per_read;
    int32_t ck_floats;      // floats per checkpoint (max over the layouts in use)
    unsigned int *counter;
    // optional indirection: the kernel works on reads list[0 .. *n_list) instead of 0 .. n_reads
    const int32_t *list;
    const int32_t *n_list;
    int32_t q_full;         // pair kernel: every listed read has exactly this query length
    // pieces: warm fronts [(read * n_warm + widx)][ck_floats] and the warm-up length
    float *warm;
    int32_t n_warm;
    int32_t warm_blocks;
    // FIX instantiations: first piece of each (read, split group) whose fronts differ (0x7fffffff: none)
    const int32_t *first_bad;   // [n_reads][n_split]
    int32_t n_split;
    const int32_t *split_first; // [n_split] piece id of piece 0
    const int32_t *split_count; // [n_split] pieces of the group
};

// ring (float2 x 128) + last-row buffer (2 values per macro-step; 2R in the generic block)
// Per-warp task state that only the per-block epilogue touches (chunk bookkeeping, running top-2, checkpoint
// schedule).  It lives in shared memory, read through a volatile pointer, so that across the unrolled macro-steps
// the registers hold nothing but the DTW state: with this state in registers ptxas ran out of room to interleave
// the two column chains (measured: -8 % on the 1 Mb shape).
enum {
    SF_ST_CLO, SF_ST_CHI, SF_ST_CHUNK, SF_ST_SI, SF_ST_LO, SF_ST_HI,
    SF_ST_EVT,    // block at whose end a checkpoint or the warm front is due (0x7fffffff: none)
    SF_ST_CK, SF_ST_PFLAGS, SF_ST_PIDX,
    SF_ST_READ,   // + half
    SF_ST_TOP = 12, // + 5 * half: s1, s2, seg, chunk, pos of the running best
    SF_ST_VALID = 22, // + half: does this half carry a read of its own (pair layout, odd lists)
    SF_ST_QLEN = 24,
    SF_ST_WORDS = 32
};
// standard DTW only: the border values of lane 0 (virtual row -1: +INF, 0 in front of a segment) for every ring
// slot, and 32 float2 of zeros that the other lanes read instead (see sf_dtw_block)
#define SF_BORDER_FLOATS (2 * SF_RING_PAIRS + 64)
__host__ __device__ inline int sf_smem_floats_per_warp(int R, bool std_dtw)
{
    return 2 * SF_RING_PAIRS + 64 * R + SF_ST_WORDS + (std_dtw ? SF_BORDER_FLOATS : 0);
}
// one wavefront checkpoint: L[R], dprev, botA per lane
__host__ __device__ inline int sf_ckpt_floats(int R) { return (R + 2) * 32; }

// register of the last query row for which the block is specialised (the generic block stores all
// R registers of the last-row lane; the specialised one stores just this one): the common full-length
// queries of the default -q values (250, 500, 100) and q = multiples of R
__host__ __device__ constexpr int sf_fast_rq(int R) { return R == 8 ? 1 : (R == 16 ? 3 : (R == 4 ? 3 : R - 1)); }

// Macro-steps unrolled together in the hot loop (body = U x (6R + 6) instructions x 16 bytes).  A fully unrolled
// block fits the 32 KB L1.5 instruction cache up to R = 8 (27 KB: measured equal to partial unrolling); beyond
// that the warps stall on instruction fetch (R = 16, 52 KB: no_instruction 5.5 stalled warps per issue, 6.9 TCUPS
// against 8.1 with an 6.5 KB body), so larger tiles unroll only as many macro-steps as fit 8 KB.
__host__ __device__ constexpr int sf_dtw_unroll(int R)
{
    if (32 * (6 * R + 6) * 16 <= 28 * 1024)
        return 32;
    int u = 16;
    while (u > 1 && u * (6 * R + 6) * 16 > 8 * 1024)
        u /= 2;
    return u;
}

__host__ __device__ constexpr int sf_dtw_min_blocks(int R) { return R <= 8 ? 10 : (R <= 12 ? 7 : (R <= 16 ? 7 : (R <= 24 ? 4 : 3))); }

// Shared-memory accesses of the hot loop go through explicit 32-bit shared addresses whose base is made
// opaque once per block: otherwise the compiler re-derives the ring / buffer address from its parts on
// every step (3-4 extra integer instructions per step) to save two registers.
__device__ __forceinline__ unsigned sf_smem_addr(const void *p)
{
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("" : "+r"(a));
    return a;
}
__device__ __forceinline__ float2 sf_lds2(unsigned addr)
{
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sf_sts2(unsigned addr, float a, float b)
{
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b));
}
__device__ __forceinline__ void sf_sts4(unsigned addr, float a, float b, float c, float d)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d));
}
// lane 0 is fed +0 (virtual row -1); an integer multiply by 0/1 keeps this off the half-rate ALU pipe
__device__ __forceinline__ float sf_mask0(float v, int nz)
{
    int ub;
    asm("mul.lo.s32 %0, %1, %2;" : "=r"(ub) : "r"(__float_as_int(v)), "r"(nz)); // IMAD: fma pipe
    return __int_as_float(ub);
}
// standard DTW: the same multiply with the border value as addend -- lane 0 (nz = 0) gets `border`, the other lanes
// read zeros there and keep the shuffled value.  One IMAD instead of FSETP + FSEL on the half-rate ALU pipe.
__device__ __forceinline__ float sf_mask0_add(float v, int nz, float border)
{
    int ub;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(ub) : "r"(__float_as_int(v)), "r"(nz), "r"(__float_as_int(border)));
    return __int_as_float(ub);
}
__device__ __forceinline__ float2 sf_border_of(const float2 v)
{
    return make_float2(v.x == SF_INF ? 0.0f : SF_INF, v.y == SF_INF ? 0.0f : SF_INF);
}

// 32 macro-steps = 64 reference columns.  RQ >= 0: the last query row sits in register RQ of lane lq and
// only its two values are handed to the chunk scan (last[2s], last[2s+1]); RQ < 0: all R registers of
// both columns are stored (last[(s*R + r)*2 + c]).
// State between macro-steps: L[r] = row r at the lane's second column, botA/botB = bottom row at the
// lane's two columns (read by the next lane one macro-step later), dprev = the `up` input of the second
// column (the diagonal of row 0 at the next macro-step's first column).
template <int R, bool STD, int RQ, int W = 32>
__device__ __forceinline__ void sf_dtw_block(const float (&x)[R], float (&L)[R], float &botA, float &botB,
                                             float &dprev, const float2 *yb, float *last, const bool is_lq,
                                             const int nz, const float2 *zb = nullptr)
{
    const unsigned full = 0xffffffffu;
    // the 32 macro-steps are unrolled U at a time (see sf_dtw_unroll)
    constexpr int U = sf_dtw_unroll(R);
    unsigned yb_o = sf_smem_addr(yb);
    unsigned last_o = sf_smem_addr(last);
    unsigned zb_o = 0; // STD: lane 0 reads its border values here, at the offsets of its reference events; others zeros
    if (STD)
        zb_o = sf_smem_addr(zb);
#pragma unroll 1
    for (int s0 = 0; s0 < 32; s0 += U) {
#pragma unroll
    for (int s = 0; s < U; s++) {
        const float2 yy = sf_lds2(yb_o + 8 * s);
        float upA = __shfl_up_sync(full, botA, 1, W);
        float upB = __shfl_up_sync(full, botB, 1, W);
        if (STD) {
            const float2 bb = sf_lds2(zb_o + 8 * s);
            upA = sf_mask0_add(upA, nz, bb.x);
            upB = sf_mask0_add(upB, nz, bb.y);
        } else {
            upA = sf_mask0(upA, nz);
            upB = sf_mask0(upB, nz);
        }
        const float next_dprev = upB;
        float dgA = dprev, dgB = upA;
        float keepA = 0.0f, keepB = 0.0f;
        float allA[RQ < 0 ? R : 1];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const float tA = x[r] - yy.x;
            const float tB = x[r] - yy.y;
            const float a = fabsf(tA) + fminf(fminf(upA, dgA), L[r]);
            const float b = fabsf(tB) + fminf(fminf(upB, dgB), a);
            dgA = L[r];
            dgB = a;
            L[r] = b;
            upA = a;
            upB = b;
            if (RQ >= 0) {
                if (r == RQ) { keepA = a; keepB = b; }
            } else {
                allA[r] = a;
            }
        }
        dprev = next_dprev;
        botA = upA;
        botB = upB;
        if (is_lq) {
            if (RQ >= 0) {
                sf_sts2(last_o + 8 * s, keepA, keepB);
            } else {
                if (R % 2 == 0) { // 16-byte stores need (s*R + r) even
#pragma unroll
                    for (int r = 0; r < R; r += 2)
                        sf_sts4(last_o + 8 * (s * R + r), allA[r], L[r], allA[r + 1], L[r + 1]);
                } else {
#pragma unroll
                    for (int r = 0; r < R; r++)
                        sf_sts2(last_o + 8 * (s * R + r), allA[r], L[r]);
                }
            }
        }
    }
        yb_o += 8 * U;
        if (STD)
            zb_o += 8 * U;
        last_o += 8 * U * (RQ >= 0 ? 1 : R);
    }
}

// do the fronts of piece `pidx` agree: the warm front it reached at its boundary against the checkpoint its
// predecessor wrote there?  n_f = floats of the layout the read ran in.  Warp-uniform result.
__device__ __forceinline__ bool sf_fronts_equal(const sf_dtw_args &a, const int read, const int pidx, const int lane, const int n_f)
{
    const sf_piece pc = a.pieces[pidx];
    const sf_group grp = a.groups[pc.gid];
    const unsigned *w = reinterpret_cast<const unsigned *>(a.warm + ((size_t)read * a.n_warm + pc.widx) * (size_t)a.ck_floats);
    const unsigned *f = reinterpret_cast<const unsigned *>(
        a.ckpt + ((size_t)read * a.ck_per_read + grp.ck_prefix + pc.b0 / grp.ck_every - 1) * (size_t)a.ck_floats);
    bool ne = false;
    for (int i = lane; i < n_f; i += 32)
        ne |= w[i] != f[i];
    return !__any_sync(0xffffffffu, ne);
}

// ---- pieces of the per-block epilogue, shared by the two layouts (W lanes per read) ----

// Task set-up of the state words (lane 0 writes, everyone syncs): chunk bookkeeping of the first segment, or of
// the piece's own columns; the block of the first checkpoint / warm front.  Returns nothing; the caller reads the
// state back through `st`.
template <bool STD, bool RESUME>
__device__ __forceinline__ void sf_task_state_init(volatile int *st, const sf_dtw_args &a, const sf_piece &pc, const sf_group &grp,
                                                   const int pidx, const int qlen, const int lq, const int lane)
{
    if (lane == 0) {
        const sf_seg seg = a.segs[grp.seg0];
        const int lo = (int)(seg.off - grp.begin);
        int hi = lo + seg.rlen;
        int chunk = 0, clo = STD ? hi - 1 : lo, ck = 0;
        if (pc.flags & 3) { // a piece of a split segment (single-segment group, never STD): the last row of its own
                            // blocks [b0, b1) is columns [64 b0 - 2 lq, 64 b1 - 2 lq)
            if (pc.flags & 2)
                hi = min(hi, SF_BLOCK_COLS * pc.b1 - 2 * lq);
            if (pc.flags & 1) {
                clo = SF_BLOCK_COLS * pc.b0 - 2 * lq;
                chunk = (clo - lo) / qlen;
                ck = pc.b0 / grp.ck_every;
            }
        }
        st[SF_ST_CLO] = clo;
        st[SF_ST_CHI] = STD ? hi : min(lo + (chunk + 1) * qlen, hi);
        st[SF_ST_CHUNK] = chunk;
        st[SF_ST_SI] = grp.seg0;
        st[SF_ST_LO] = lo;
        st[SF_ST_HI] = hi;
        st[SF_ST_CK] = ck;
        // bit 0: has a predecessor, bit 1: has a successor, bit 2: the head edge is still to come
        st[SF_ST_PFLAGS] = pc.flags | ((pc.flags & 1) << 2);
        st[SF_ST_PIDX] = pidx;
        int evt = 0x7fffffff;
        if (grp.ck_every > 0 && ck < grp.n_ck)
            evt = (ck + 1) * grp.ck_every - 1;
        if (!RESUME && (pc.flags & 1))
            evt = pc.b0 - 1; // the warm front comes first
        st[SF_ST_EVT] = evt;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            st[SF_ST_TOP + 5 * h + 0] = __float_as_int(SF_INF);
            st[SF_ST_TOP + 5 * h + 1] = __float_as_int(SF_INF);
            st[SF_ST_TOP + 5 * h + 2] = -1;
            st[SF_ST_TOP + 5 * h + 3] = 0;
            st[SF_ST_TOP + 5 * h + 4] = -1;
        }
    }
    __syncwarp();
}

// A chunk is complete: (m, mp) = its first strict minimum for this lane's read (already reduced over the read's
// lanes).  Books it -- as the head / tail edge of a piece or into the running top-2 of read `half` -- and opens the
// next chunk.  `writer`: the one lane of the read that stores.
template <bool STD>
__device__ __forceinline__ void sf_chunk_done(volatile int *st, const sf_dtw_args &a, const float m, const int mp, const int half,
                                              const bool writer, const int lane)
{
    const int pidx = st[SF_ST_PIDX];
    const bool valid = st[SF_ST_VALID + half] != 0;
    const int qlen = st[SF_ST_QLEN];
    const int pflags = st[SF_ST_PFLAGS];
    int chunk = st[SF_ST_CHUNK], si = st[SF_ST_SI], lo = st[SF_ST_LO], hi = st[SF_ST_HI];
    const int chi = st[SF_ST_CHI];
    const int pos = mp >= 0 ? mp - lo : -1;
    if (pflags & 4) { // the chunk cut by the piece's start: reported apart (head edge)
        if (writer && valid) {
            sf_taskres *out = a.res + (size_t)st[SF_ST_READ + half] * a.n_pieces + pidx;
            out->hmin = m; out->hpos = pos; out->hchunk = chunk;
        }
    } else if ((pflags & 2) && chi >= hi) { // the chunk cut by the piece's end (tail edge)
        if (writer && valid) {
            sf_taskres *out = a.res + (size_t)st[SF_ST_READ + half] * a.n_pieces + pidx;
            out->tmin = m; out->tpos = pos; out->tchunk = chunk;
        }
    } else if (writer) { // update_aln(): a later candidate with an equal score ranks better
        volatile int *t = st + SF_ST_TOP + 5 * half;
        const float s1 = __int_as_float(t[0]), s2 = __int_as_float(t[1]);
        if (m <= s1) {
            t[1] = __float_as_int(s1); t[0] = __float_as_int(m); t[2] = si; t[3] = chunk; t[4] = pos;
        } else if (m < s2) {
            t[1] = __float_as_int(m);
        }
    }
    __syncwarp();
    if (lane == 0) {
        int clo = chi;
        chunk++;
        if (clo >= hi) {
            const sf_group grp = a.groups[a.pieces[pidx].gid];
            si++;
            chunk = 0;
            if (si < grp.seg0 + grp.nseg) {
                const sf_seg seg = a.segs[si];
                lo = (int)(seg.off - grp.begin);
                hi = lo + seg.rlen;
                clo = STD ? hi - 1 : lo;
            } else {
                clo = 0x7fffffff;
                hi = 0x7fffffff;
            }
        }
        st[SF_ST_CLO] = clo;
        st[SF_ST_CHI] = (clo == 0x7fffffff) ? 0x7fffffff : (STD ? hi : min(clo + qlen, hi));
        st[SF_ST_CHUNK] = chunk;
        st[SF_ST_SI] = si;
        st[SF_ST_LO] = lo;
        st[SF_ST_HI] = hi;
        st[SF_ST_PFLAGS] = pflags & ~4;
    }
    __syncwarp();
}

// End of block b == st[SF_ST_EVT]: write the regular checkpoint and / or the piece's warm front (layout
// [R+2][W]: L[r], dprev, botA per lane), then schedule the next one.
template <int R, int W, bool RESUME>
__device__ __forceinline__ void sf_block_event(volatile int *st, const sf_dtw_args &a, const int b, const int half,
                                               const int ll, const int lane, const float (&L)[R], const float dprev, const float botA)
{
    const int read = st[SF_ST_READ + half];
    const bool valid = st[SF_ST_VALID + half] != 0;
    const sf_piece pc = a.pieces[st[SF_ST_PIDX]];
    const sf_group grp = a.groups[pc.gid];
    int ck = st[SF_ST_CK];
    float *c = nullptr;
    if (grp.ck_every > 0 && ck < grp.n_ck && (b + 1) == (ck + 1) * grp.ck_every) {
        c = a.ckpt + ((size_t)read * a.ck_per_read + grp.ck_prefix + ck) * (size_t)a.ck_floats;
        ck++;
    } else if (!RESUME && (pc.flags & 1) && b + 1 == pc.b0) {
        c = a.warm + ((size_t)read * a.n_warm + pc.widx) * (size_t)a.ck_floats;
    }
    if (c && valid) {
#pragma unroll
        for (int r = 0; r < R; r++)
            c[r * W + ll] = L[r];
        c[R * W + ll] = dprev;
        c[(R + 1) * W + ll] = botA;
    }
    __syncwarp();
    if (lane == 0) {
        st[SF_ST_CK] = ck;
        st[SF_ST_EVT] = (grp.ck_every > 0 && ck < grp.n_ck) ? (ck + 1) * grp.ck_every - 1 : 0x7fffffff;
    }
    __syncwarp();
}

// One (read, piece) task of the warp-per-read layout.  RESUME: start at the piece's boundary from the checkpoint the
// predecessor left there (redo of a piece whose warm front did not verify) instead of warming up.
template <int R, bool STD, bool RESUME>
__device__ __forceinline__ void sf_score_task(const sf_dtw_args &a, const int read, const int pidx, const int lane,
                                              float2 *ring, float *last)
{
    const unsigned full = 0xffffffffu;
    volatile int *st = reinterpret_cast<volatile int *>(last + 64 * R);
    float2 *bring = reinterpret_cast<float2 *>(last + 64 * R + SF_ST_WORDS); // STD only
    const sf_piece pc = a.pieces[pidx];
    const sf_group grp = a.groups[pc.gid];
    const int qlen = a.info[read].qlen;
    if (qlen <= 0) {
        if (lane == 0) {
            sf_taskres *out = a.res + (size_t)read * a.n_pieces + pidx;
            out->s1 = SF_INF; out->s2 = SF_INF; out->seg = -1; out->chunk = 0; out->pos = -1;
            out->hmin = SF_INF; out->hpos = -1; out->hchunk = -1; out->tmin = SF_INF; out->tpos = -1; out->tchunk = -2;
        }
        return;
    }
    const int lq = (qlen - 1) / R; // lane holding the last query row
    const int rq = (qlen - 1) % R; // its register
    const bool is_lq = lane == lq;
    const bool fast = rq == sf_fast_rq(R); // warp-uniform
    const int nz = lane != 0;

    const float *y = a.stream + grp.begin;
    const int n_pos = (int)(grp.end - grp.begin); // includes the leading sentinel
    // the last column (n_pos-1) is in pair (n_pos-1)/2 and reaches lane lq lq macro-steps later
    const int b_end = (pc.flags & 2) ? pc.b1 : ((n_pos - 1) / 2 + lq + 32) >> 5;
    const int b_first = RESUME ? pc.b0 : ((pc.flags & 1) ? max(0, pc.b0 - a.warm_blocks) : 0);

    // query rows of this lane; rows past qlen are padding (finite, never read back)
    float x[R], L[R];
    const float *q = a.queries + (size_t)read * a.q_cap;
    float botA = SF_INF, botB = SF_INF;
    float dprev = (lane == 0 && !STD) ? 0.0f : SF_INF;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int row = lane * R + r;
        x[r] = row < qlen ? q[row] : 0.0f;
        L[r] = SF_INF;
    }
    if (RESUME) {
        const float *c = a.ckpt + ((size_t)read * a.ck_per_read + grp.ck_prefix + pc.b0 / grp.ck_every - 1) * (size_t)a.ck_floats;
#pragma unroll
        for (int r = 0; r < R; r++)
            L[r] = c[r * 32 + lane];
        dprev = c[R * 32 + lane];
        botA = c[(R + 1) * 32 + lane];
        botB = L[R - 1];
    }

    __syncwarp();
    // ring: block b (pairs 32b .. 32b+31) lives at slots (b&1)*32 + j and again 64 slots later; the block "before"
    // the first one reads as +INF columns, and so does the first column of a warm-up (the virtual column -1 of the
    // windowed recurrence); a resumed task continues inside the segment and needs the real previous block
    {
        const int c0 = SF_BLOCK_COLS * b_first + 2 * lane, c1 = c0 + 1;
        float2 v = make_float2(c0 < n_pos ? y[c0] : SF_INF, c1 < n_pos ? y[c1] : SF_INF);
        if (!RESUME && b_first > 0 && lane == 0)
            v.x = SF_INF;
        float2 pv = make_float2(SF_INF, SF_INF);
        if (RESUME)
            pv = make_float2(c0 - SF_BLOCK_COLS < n_pos ? y[c0 - SF_BLOCK_COLS] : SF_INF,
                             c1 - SF_BLOCK_COLS < n_pos ? y[c1 - SF_BLOCK_COLS] : SF_INF);
        const int cur = (b_first & 1) * 32 + lane, prv = ((b_first & 1) ^ 1) * 32 + lane;
        ring[cur] = v; ring[cur + 64] = v;
        ring[prv] = pv; ring[prv + 64] = pv;
        if (STD) {
            bring[cur] = bring[cur + 64] = sf_border_of(v);
            bring[prv] = bring[prv + 64] = sf_border_of(pv);
            bring[SF_RING_PAIRS + lane] = make_float2(0.0f, 0.0f);
        }
    }
    if (lane == 0) {
        st[SF_ST_READ] = read;
        st[SF_ST_VALID] = 1;
        st[SF_ST_QLEN] = qlen;
    }
    sf_task_state_init<STD, RESUME>(st, a, pc, grp, pidx, qlen, lq, lane);

    float rmin = SF_INF; // per lane running minimum of the open chunk
    int rpos = -1;

    for (int b = b_first; b < b_end; b++) {
        // prefetch the next 64 reference events; published after the 32 macro-steps below
        const int nidx = SF_BLOCK_COLS * (b + 1) + 2 * lane;
        const float yn0 = nidx < n_pos ? __ldg(y + nidx) : SF_INF;
        const float yn1 = nidx + 1 < n_pos ? __ldg(y + nidx + 1) : SF_INF;
        const float2 *yb = ring + ((b & 1) ? 32 : 64) - lane;
        const float2 *zb = lane == 0 ? bring + ((b & 1) ? 32 : 64) : bring + SF_RING_PAIRS;

        if (fast)
            sf_dtw_block<R, STD, sf_fast_rq(R)>(x, L, botA, botB, dprev, yb, last, is_lq, nz, zb);
        else
            sf_dtw_block<R, STD, -1>(x, L, botA, botB, dprev, yb, last, is_lq, nz, zb);
        __syncwarp();

        // ---- last-row chunk minima (sigfish.c:891-901): this block produced the last row of columns
        //      p0 .. p0+63; lane l looks at p0+2l and p0+2l+1 ----
        {
            const int p0 = SF_BLOCK_COLS * b - 2 * lq;
            const int pos0 = p0 + 2 * lane;
            float v0, v1;
            if (fast) {
                const float2 v = *reinterpret_cast<const float2 *>(last + 2 * lane);
                v0 = v.x; v1 = v.y;
            } else {
                const float2 v = *reinterpret_cast<const float2 *>(last + (lane * R + rq) * 2);
                v0 = v.x; v1 = v.y;
            }
            for (;;) {
                const int clo = st[SF_ST_CLO], chi = st[SF_ST_CHI];
                if (pos0 >= clo && pos0 < chi && v0 < rmin) {
                    rmin = v0;
                    rpos = pos0;
                }
                if (pos0 + 1 >= clo && pos0 + 1 < chi && v1 < rmin) {
                    rmin = v1;
                    rpos = pos0 + 1;
                }
                if (chi > p0 + SF_BLOCK_COLS)
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
                sf_chunk_done<STD>(st, a, m, mp, 0, lane == 0, lane);
                rmin = SF_INF;
                rpos = -1;
            }
        }

        // ---- checkpoint of the skewed wavefront (for the start-coordinate pass; the one at a piece's end is also
        //      what the next piece's warm front is verified against), warm front of a piece ----
        if (b == st[SF_ST_EVT])
            sf_block_event<R, 32, RESUME>(st, a, b, 0, lane, lane, L, dprev, botA);

        // publish block b+1 of the reference events (overwrites block b-1)
        {
            const int slot = ((b + 1) & 1) * 32 + lane;
            const float2 v = make_float2(yn0, yn1);
            ring[slot] = v;
            ring[slot + 64] = v;
            if (STD)
                bring[slot] = bring[slot + 64] = sf_border_of(v);
        }
        __syncwarp();
    }

    if (lane == 0) {
        sf_taskres *out = a.res + (size_t)st[SF_ST_READ] * a.n_pieces + st[SF_ST_PIDX];
        volatile int *t = st + SF_ST_TOP;
        out->s1 = __int_as_float(t[0]); out->s2 = __int_as_float(t[1]); out->seg = t[2]; out->chunk = t[3]; out->pos = t[4];
    }
    __syncwarp();
}

// FIX: the redo pass.  Walks the pieces of every (read, split group) that sf_verify_kernel flagged, from the first
// piece whose fronts differ: that piece is recomputed from its predecessor's front (which is the full-matrix state
// by induction), which rewrites the checkpoint at its own end; every later piece is then checked against the
// front now standing before it and recomputed if it differs.
template <int R, bool STD, bool FIX = false>
__global__ void __launch_bounds__(SF_DTW_THREADS, sf_dtw_min_blocks(R)) sf_dtw_score_kernel(const sf_dtw_args a)
{
    extern __shared__ float2 sf_smem2[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float2 *ring = sf_smem2 + warp * (sf_smem_floats_per_warp(R, STD) / 2);
    float *last = reinterpret_cast<float *>(ring + SF_RING_PAIRS);
    const unsigned full = 0xffffffffu;
    const int n_list = a.list ? *a.n_list : a.n_reads;
    const unsigned n_tasks = (unsigned)(FIX ? a.n_split : a.n_pieces) * (unsigned)n_list;
    // a kernel launched behind this one with programmatic stream serialisation (the pair kernel: it shares no data
    // with this one) may start as soon as every block of this grid is resident
    if (!FIX)
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    for (;;) {
        unsigned task = 0;
        if (lane == 0)
            task = atomicAdd(a.counter, 1u);
        task = __shfl_sync(full, task, 0);
        if (task >= n_tasks)
            break;
        const int gi = task / (unsigned)n_list;
        const int li = task - gi * n_list;
        const int read = a.list ? a.list[li] : li;
        if (!FIX) {
            sf_score_task<R, STD, false>(a, read, a.order[gi], lane, ring, last);
        } else {
            const int k0 = a.first_bad[(size_t)read * a.n_split + gi];
            if (k0 == 0x7fffffff)
                continue;
            const int p0 = a.split_first[gi], K = a.split_count[gi];
#pragma unroll 1
            for (int k = k0; k < K; k++) {
                if (k > k0 && sf_fronts_equal(a, read, p0 + k, lane, sf_ckpt_floats(R)))
                    continue;
                sf_score_task<R, STD, true>(a, read, p0 + k, lane, ring, last);
                __syncwarp();
            }
        }
    }
}


// ---- two full-length reads per warp ---------------------------------------------------------------------
// For q <= 256 a read needs only 16 lanes when each lane holds 16 rows, and the per-macro-step overheads (LDS,
// shuffles, border IMADs, last-row store) are then shared by twice the cells (measured on a 1 Mb contig: 8.3 TCUPS
// against 7.7 at R = 8).  Lanes 0-15 carry read list[2u], lanes 16-31 read list[2u+1]; both stream the same
// segment group, so the ring and all chunk bookkeeping are shared and only the minima are kept per read.  Every
// listed read has exactly q_full events (sf_partition_kernel), i.e. the last query row is register RQ of lane
// (q_full-1)/R of its half; shorter queries go through sf_dtw_score_kernel, which is launched first and releases
// this kernel as soon as its blocks are resident (see run_stages in sfgpu.cu).  Checkpoints use the half-warp
// layout [R+2][16] (consumed by sf_trace_start<R, STD, 16>).
#define SF_PAIR_LANES 16
// last-row staging of one read: 64 floats + 4 of padding, so that the two storing lanes hit different banks
#define SF_PAIR_LAST 68
__host__ __device__ inline int sf_pair_smem_floats_per_warp(bool std_dtw)
{
    return 2 * SF_RING_PAIRS + 2 * SF_PAIR_LAST + SF_ST_WORDS + (std_dtw ? SF_BORDER_FLOATS : 0);
}

// One (pair of reads, piece) task.  `read` / `valid` are per half; RESUME as in sf_score_task.
template <int R, bool STD, int RQ, bool RESUME>
__device__ __forceinline__ void sf_pair_task(const sf_dtw_args &a, const int read, const bool valid, const int pidx,
                                             const int lane, float2 *ring, float *last)
{
    constexpr int W = SF_PAIR_LANES;
    const unsigned full = 0xffffffffu;
    volatile int *st = reinterpret_cast<volatile int *>(last + 2 * SF_PAIR_LAST);
    float2 *bring = reinterpret_cast<float2 *>(last + 2 * SF_PAIR_LAST + SF_ST_WORDS); // STD only
    const int ll = lane & (W - 1);  // lane inside the read
    const int half = lane / W;      // which read of the pair
    const int qlen = a.q_full;
    const int lq = (qlen - 1) / R;  // lane (inside the half) holding the last query row; its register is RQ
    const bool is_lq = ll == lq;
    const int nz = ll != 0;
    const sf_piece pc = a.pieces[pidx];
    const sf_group grp = a.groups[pc.gid];

    const float *y = a.stream + grp.begin;
    const int n_pos = (int)(grp.end - grp.begin); // includes the leading sentinel
    const int b_end = (pc.flags & 2) ? pc.b1 : ((n_pos - 1) / 2 + lq + 32) >> 5;
    const int b_first = RESUME ? pc.b0 : ((pc.flags & 1) ? max(0, pc.b0 - a.warm_blocks) : 0);

    float x[R], L[R];
    const float *q = a.queries + (size_t)read * a.q_cap;
    float botA = SF_INF, botB = SF_INF;
    float dprev = (ll == 0 && !STD) ? 0.0f : SF_INF;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int row = ll * R + r;
        x[r] = row < qlen ? q[row] : 0.0f;
        L[r] = SF_INF;
    }
    if (RESUME) {
        const float *c = a.ckpt + ((size_t)read * a.ck_per_read + grp.ck_prefix + pc.b0 / grp.ck_every - 1) * (size_t)a.ck_floats;
#pragma unroll
        for (int r = 0; r < R; r++)
            L[r] = c[r * W + ll];
        dprev = c[R * W + ll];
        botA = c[(R + 1) * W + ll];
        botB = L[R - 1];
    }

    __syncwarp();
    {
        const int c0 = SF_BLOCK_COLS * b_first + 2 * lane, c1 = c0 + 1;
        float2 v = make_float2(c0 < n_pos ? y[c0] : SF_INF, c1 < n_pos ? y[c1] : SF_INF);
        if (!RESUME && b_first > 0 && lane == 0)
            v.x = SF_INF; // virtual column -1 of the warm-up window
        float2 pv = make_float2(SF_INF, SF_INF);
        if (RESUME)
            pv = make_float2(c0 - SF_BLOCK_COLS < n_pos ? y[c0 - SF_BLOCK_COLS] : SF_INF,
                             c1 - SF_BLOCK_COLS < n_pos ? y[c1 - SF_BLOCK_COLS] : SF_INF);
        const int cur = (b_first & 1) * 32 + lane, prv = ((b_first & 1) ^ 1) * 32 + lane;
        ring[cur] = v; ring[cur + 64] = v;
        ring[prv] = pv; ring[prv + 64] = pv;
        if (STD) {
            bring[cur] = bring[cur + 64] = sf_border_of(v);
            bring[prv] = bring[prv + 64] = sf_border_of(pv);
            bring[SF_RING_PAIRS + lane] = make_float2(0.0f, 0.0f);
        }
    }
    // chunk bookkeeping is shared by the two reads (same query length, same segments); the running best is per read
    if (ll == 0) {
        st[SF_ST_READ + half] = read;
        st[SF_ST_VALID + half] = valid ? 1 : 0;
        st[SF_ST_QLEN] = qlen;
    }
    sf_task_state_init<STD, RESUME>(st, a, pc, grp, pidx, qlen, lq, lane);

    float rmin = SF_INF;
    int rpos = -1;

    for (int b = b_first; b < b_end; b++) {
        const int nidx = SF_BLOCK_COLS * (b + 1) + 2 * lane;
        const float yn0 = nidx < n_pos ? __ldg(y + nidx) : SF_INF;
        const float yn1 = nidx + 1 < n_pos ? __ldg(y + nidx + 1) : SF_INF;
        const float2 *yb = ring + ((b & 1) ? 32 : 64) - ll;
        const float2 *zb = ll == 0 ? bring + ((b & 1) ? 32 : 64) : bring + SF_RING_PAIRS;

        sf_dtw_block<R, STD, RQ, W>(x, L, botA, botB, dprev, yb, last + SF_PAIR_LAST * half, is_lq, nz, zb);
        __syncwarp();

        // ---- last-row chunk minima: this block produced columns p0 .. p0+63 of both reads; lane ll of a
        //      half looks at p0+4ll .. p0+4ll+3 of its own read ----
        {
            const int p0 = SF_BLOCK_COLS * b - 2 * lq;
            const int pos0 = p0 + 4 * ll;
            const float4 v = *reinterpret_cast<const float4 *>(last + SF_PAIR_LAST * half + 4 * ll);
            for (;;) {
                const int clo = st[SF_ST_CLO], chi = st[SF_ST_CHI];
                if (pos0 >= clo && pos0 < chi && v.x < rmin) { rmin = v.x; rpos = pos0; }
                if (pos0 + 1 >= clo && pos0 + 1 < chi && v.y < rmin) { rmin = v.y; rpos = pos0 + 1; }
                if (pos0 + 2 >= clo && pos0 + 2 < chi && v.z < rmin) { rmin = v.z; rpos = pos0 + 2; }
                if (pos0 + 3 >= clo && pos0 + 3 < chi && v.w < rmin) { rmin = v.w; rpos = pos0 + 3; }
                if (chi > p0 + SF_BLOCK_COLS)
                    break;
                float m = rmin;
                int mp = rpos;
#pragma unroll
                for (int o = W / 2; o > 0; o >>= 1) { // stays inside the half
                    const float om = __shfl_xor_sync(full, m, o);
                    const int op = __shfl_xor_sync(full, mp, o);
                    if (om < m || (om == m && (unsigned)op < (unsigned)mp)) {
                        m = om;
                        mp = op;
                    }
                }
                sf_chunk_done<STD>(st, a, m, mp, half, ll == 0, lane);
                rmin = SF_INF;
                rpos = -1;
            }
        }

        // ---- checkpoint (half-warp layout, one per read) / warm front of a piece ----
        if (b == st[SF_ST_EVT])
            sf_block_event<R, W, RESUME>(st, a, b, half, ll, lane, L, dprev, botA);

        {
            const int slot = ((b + 1) & 1) * 32 + lane;
            const float2 v = make_float2(yn0, yn1);
            ring[slot] = v;
            ring[slot + 64] = v;
            if (STD)
                bring[slot] = bring[slot + 64] = sf_border_of(v);
        }
        __syncwarp();
    }

    if (ll == 0 && st[SF_ST_VALID + half]) {
        sf_taskres *out = a.res + (size_t)st[SF_ST_READ + half] * a.n_pieces + st[SF_ST_PIDX];
        volatile int *t = st + SF_ST_TOP + 5 * half;
        out->s1 = __int_as_float(t[0]); out->s2 = __int_as_float(t[1]); out->seg = t[2]; out->chunk = t[3]; out->pos = t[4];
    }
    __syncwarp();
}

// resident blocks per SM of the pair kernel (measured on the 1 Mb shape, profiles/r02_pair_rows_follow_q.txt and
// r02_pair_blocks_per_sm_ab.txt: 8 blocks of 64 registers against 7 of 72 are +1.9 % at R = 10, +2.7 % at R = 12, +1.9 %
// at R = 13, but -3.6 % at R = 9, -2.2 % at R = 15 and -1.7 % at R = 16; 6 blocks of 80 registers: +0.7 % at R = 16 (inside
// the box-to-box noise), -2.7 % at R = 15)
__host__ __device__ constexpr int sf_pair_min_blocks(int R) { return R <= 8 ? 10 : (R == 9 ? 7 : (R <= 13 ? 8 : 7)); }

template <int R, bool STD, int RQ, bool FIX = false>
__global__ void __launch_bounds__(SF_DTW_THREADS, sf_pair_min_blocks(R)) sf_dtw_pair_kernel(const sf_dtw_args a)
{
    constexpr int W = SF_PAIR_LANES;
    extern __shared__ float2 sf_smem2[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int half = lane / W;
    float2 *ring = sf_smem2 + warp * (sf_pair_smem_floats_per_warp(STD) / 2);
    float *last = reinterpret_cast<float *>(ring + SF_RING_PAIRS);
    const unsigned full = 0xffffffffu;
    const int n_list = *a.n_list;
    // FIX: one read per warp (both halves carry it, the upper one writes nothing)
    const int n_units = FIX ? n_list : (n_list + 1) / 2;
    const unsigned n_tasks = (unsigned)(FIX ? a.n_split : a.n_pieces) * (unsigned)n_units;

    for (;;) {
        unsigned task = 0;
        if (lane == 0)
            task = atomicAdd(a.counter, 1u);
        task = __shfl_sync(full, task, 0);
        if (task >= n_tasks)
            break;
        const int gi = task / (unsigned)n_units;
        const int unit = task - gi * n_units;
        if (!FIX) {
            const int idx = 2 * unit + half;
            const bool valid = idx < n_list;           // an odd list: the second half repeats the first read
            const int read = a.list[valid ? idx : 2 * unit];
            sf_pair_task<R, STD, RQ, false>(a, read, valid, a.order[gi], lane, ring, last);
        } else {
            const int read = a.list[unit];
            const int k0 = a.first_bad[(size_t)read * a.n_split + gi];
            if (k0 == 0x7fffffff)
                continue;
            const int p0 = a.split_first[gi], K = a.split_count[gi];
#pragma unroll 1
            for (int k = k0; k < K; k++) {
                if (k > k0 && sf_fronts_equal(a, read, p0 + k, lane, (R + 2) * W))
                    continue;
                sf_pair_task<R, STD, RQ, true>(a, read, half == 0, p0 + k, lane, ring, last);
                __syncwarp();
            }
        }
    }
    // This grid was released early by sf_dtw_score_kernel (griddepcontrol.launch_dependents).  The kernels behind it
    // in the stream read the results of BOTH; they are ordered behind this grid only, so this grid must not
    // complete before its prerequisite has completed and flushed: wait for it here, after the work.
    if (!FIX)
        asm volatile("griddepcontrol.wait;" ::: "memory");
}

#ifndef SF_PAIR_INST_IMPL // the non-template kernels are compiled by sfgpu.cu only
// Marks, for every (read, split group), the first piece whose warm front differs from the checkpoint its predecessor
// wrote at the same boundary (atomicMin into first_bad, preset to 0x7fffffff).  One warp per (read, warm front).
struct sf_verify_args {
    const sf_piece *pieces;
    const sf_group *groups;
    const int32_t *warm_piece;  // [n_warm] piece id of warm front w
    const sf_readinfo *info;
    const float *warm, *ckpt;
    int32_t n_reads, n_warm, n_split, ck_floats;
    int64_t ck_per_read;
    int32_t n_f, n_f_pair;      // floats of the warp-per-read layout / of the pair layout (reads with status bit 5)
    int32_t *first_bad;
    int32_t *n_bad;             // statistics: fronts that differed in this batch
};

__global__ void sf_verify_kernel(const sf_verify_args a)
{
    const int lane = threadIdx.x & 31;
    const long long n_items = (long long)a.n_reads * a.n_warm;
    for (long long item = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < n_items;
         item += ((long long)gridDim.x * blockDim.x) >> 5) {
        const int read = (int)(item / a.n_warm);
        const int w = (int)(item - (long long)read * a.n_warm);
        const sf_readinfo ri = a.info[read];
        if (ri.qlen <= 0)
            continue;
        const sf_piece pc = a.pieces[a.warm_piece[w]];
        const sf_group grp = a.groups[pc.gid];
        const unsigned *x = reinterpret_cast<const unsigned *>(a.warm + ((size_t)read * a.n_warm + w) * (size_t)a.ck_floats);
        const unsigned *f = reinterpret_cast<const unsigned *>(
            a.ckpt + ((size_t)read * a.ck_per_read + grp.ck_prefix + pc.b0 / grp.ck_every - 1) * (size_t)a.ck_floats);
        const int n_f = (ri.status & 32) ? a.n_f_pair : a.n_f;
        bool ne = false;
        for (int i = lane; i < n_f; i += 32)
            ne |= x[i] != f[i];
        if (__any_sync(0xffffffffu, ne) && lane == 0) {
            atomicMin(a.first_bad + (size_t)read * a.n_split + pc.sidx, pc.k);
            atomicAdd(a.n_bad, 1);
        }
    }
}

// Splits the reads of a batch into those with exactly q_full events (pair kernel; flagged with status bit 5)
// and the rest (warp-per-read kernel), both in ascending read order.  One block of SF_PART_THREADS threads.
#define SF_PART_THREADS 1024
__global__ void __launch_bounds__(SF_PART_THREADS) sf_partition_kernel(sf_readinfo *info, int n_reads, int q_full, int32_t *list_full,
                                                                        int32_t *list_other, int32_t *counts)
{
    __shared__ int wf[SF_PART_THREADS / 32], wo[SF_PART_THREADS / 32];
    __shared__ int base_f, base_o;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        base_f = 0;
        base_o = 0;
    }
    __syncthreads();
    for (int base = 0; base < n_reads; base += SF_PART_THREADS) {
        const int i = base + tid;
        const int ql = i < n_reads ? info[i].qlen : -1;
        const bool isf = ql == q_full, iso = ql > 0 && !isf;
        const unsigned mf = __ballot_sync(0xffffffffu, isf), mo = __ballot_sync(0xffffffffu, iso);
        if (lane == 0) {
            wf[warp] = __popc(mf);
            wo[warp] = __popc(mo);
        }
        __syncthreads();
        int pf = base_f, po = base_o, tf = 0, to = 0;
        for (int w = 0; w < SF_PART_THREADS / 32; w++) {
            if (w < warp) {
                pf += wf[w];
                po += wo[w];
            }
            tf += wf[w];
            to += wo[w];
        }
        const unsigned below = (1u << lane) - 1u;
        if (isf) {
            list_full[pf + __popc(mf & below)] = i;
            info[i].status |= 32;
        }
        if (iso)
            list_other[po + __popc(mo & below)] = i;
        __syncthreads();
        if (tid == 0) {
            base_f += tf;
            base_o += to;
        }
        __syncthreads();
    }
    if (tid == 0) {
        counts[0] = base_f;
        counts[1] = base_o;
    }
}
#endif
