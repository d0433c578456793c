// sf_trace.cuh -- per-read merge of the task results and the start-coordinate pass.
//
// Replaces (reference, paths relative to /root/reference):
//   src/sigfish.c:575-626 update_aln()      -- ordering of candidates across refs/strands/chunks
//   src/sigfish.c:969-970                   -- score = list[4], score2 = list[3]
//   src/cdtw.c:98-167 path(), 192-227 subsequence_path() -- reduced to what the caller reads:
//                                              pos_st = p.py[0] of the winning hit
//
// The reference back-tracks through the stored qlen x rlen cost matrix.  Here no matrix exists;
// instead the winning (segment, end column) is recomputed over a short window with *forward
// start-pointer propagation*: every cell carries the reference column at which its optimal path
// left row 0, inherited from the predecessor chosen by the same rule as cdtw.c:134-146
// (diagonal if it equals the minimum, else left, else up); start(0,j) = j.  SURVEY.md F4 shows
// the equivalence.  The window starts either at the segment's sentinel column (all-INF state,
// no history needed) or at a wavefront checkpoint written by the score kernel; cells on the
// restart front carry an "exit code" instead of a column, and if the path of the target cell
// leaves the window through the front the pass is repeated from an earlier restart point with
// that front cell as the new target.  The pass runs the score kernel's macro-steps (two columns per
// lane per step) so that its checkpoints can be resumed as they are.
#pragma once
#include <cuda_runtime.h>
#include "sf_types.cuh"

struct sf_trace_args {
    const float *stream;
    const sf_seg *segs;
    const sf_group *groups;
    const int32_t *seg_group; // segment -> group id
    int32_t n_groups;
    const sf_piece *pieces;   // the DTW tasks of one read (split level of the batch)
    int32_t n_pieces;
    int32_t n_reads;
    const float *queries;
    const sf_readinfo *info;
    int32_t q_cap;
    const sf_taskres *res;    // [n_reads][n_pieces]
    const float *ckpt;
    int64_t ck_per_read;
    int32_t ck_floats;        // floats per checkpoint
    sf_hit *hits;
    int32_t min_window;       // restart at least this many columns before the target
    // paired reads (sf_partition_kernel): traced two per warp by sf_trace_pair_kernel
    const int32_t *list_full;
    const int32_t *n_full;
};

// a == b ? x : y as one FSETP + SEL (the compiler otherwise turns the backtrack rule's nested selection into a
// branch per row)
__device__ __forceinline__ int sf_sel_eq(float a, float b, int x, int y)
{
    int r;
    asm("{\n\t.reg .pred p;\n\tsetp.eq.f32 p, %1, %2;\n\tselp.s32 %0, %3, %4, p;\n\t}" : "=r"(r) : "f"(a), "f"(b), "r"(x), "r"(y));
    return r;
}

struct sf_true { static constexpr bool value = true; };
struct sf_false { static constexpr bool value = false; };

struct sf_top {
    float s1, s2;
    int seg, chunk, pos;
};

// is candidate (s, seg, chunk) ranked better than (t, tseg, tchunk)?  lower score first; of equal
// scores the one processed later wins (sigfish.c:577-583)
__device__ __forceinline__ bool sf_better(float s, int seg, int chunk, float t, int tseg, int tchunk)
{
    if (s < t) return true;
    if (s > t) return false;
    if (seg != tseg) return seg > tseg;
    return chunk > tchunk;
}

__device__ __forceinline__ void sf_top_merge(sf_top &a, const sf_top &b)
{
    if (b.seg >= 0 && (a.seg < 0 || sf_better(b.s1, b.seg, b.chunk, a.s1, a.seg, a.chunk))) {
        const float second = fminf(a.seg >= 0 ? a.s1 : SF_INF, b.s2);
        a.s1 = b.s1; a.seg = b.seg; a.chunk = b.chunk; a.pos = b.pos;
        a.s2 = fminf(second, a.s2);
    } else {
        a.s2 = fminf(a.s2, fminf(b.seg >= 0 ? b.s1 : SF_INF, b.s2));
    }
}

// Merges the results of task t of a read into `top`: the candidates of the chunks lying wholly inside the task,
// and -- for a piece with a successor -- the chunk cut by the boundary between the two: its minimum is the smaller
// of the two parts, at equal values the earlier column, i.e. the part in this piece (first strict minimum,
// sigfish.c:891-901).  When the boundary happens to fall between two chunks the two parts are candidates of their own.
__device__ __forceinline__ void sf_merge_task(sf_top &top, const sf_trace_args &a, const sf_taskres *res, const int t)
{
    const sf_taskres tr = res[t];
    sf_top b; b.s1 = tr.s1; b.s2 = tr.s2; b.seg = tr.seg; b.chunk = tr.chunk; b.pos = tr.pos;
    sf_top_merge(top, b);
    const sf_piece pc = a.pieces[t];
    if (pc.flags & 2) {
        const sf_taskres nx = res[t + 1]; // the next piece of the same segment
        const int seg = a.groups[pc.gid].seg0;
        sf_top e; e.s2 = SF_INF; e.seg = seg;
        if (tr.tchunk == nx.hchunk) {
            const bool first = tr.tmin <= nx.hmin;
            e.s1 = first ? tr.tmin : nx.hmin; e.chunk = tr.tchunk; e.pos = first ? tr.tpos : nx.hpos;
            sf_top_merge(top, e);
        } else {
            e.s1 = tr.tmin; e.chunk = tr.tchunk; e.pos = tr.tpos;
            sf_top_merge(top, e);
            e.s1 = nx.hmin; e.chunk = nx.hchunk; e.pos = nx.hpos;
            sf_top_merge(top, e);
        }
    }
}

// Start coordinate of the winning cell (qlen-1, top_pos of segment `seg`) for one read; W lanes hold the read
// (32: one read per warp; 16: the half-warp layout of sf_dtw_pair_kernel, whose checkpoints are [R+2][16]).
// With W = 16 the upper half of the warp mirrors the lower half.
template <int R, bool STD, int W>
__device__ __forceinline__ int sf_trace_start(const sf_trace_args &a, const int read, const int lane_in_warp, const int qlen,
                                              const int top_pos, const sf_seg &seg, const sf_group &grp)
{
    const unsigned full = 0xffffffffu;
    const int lane = lane_in_warp & (W - 1);
    const float *y = a.stream + grp.begin;          // position 0 = the group's leading sentinel
    const int n_pos = (int)(grp.end - grp.begin);
    const int seg_lo = (int)(seg.off - grp.begin);  // position of the segment's column 0
    float x[R];
    const float *q = a.queries + (size_t)read * a.q_cap;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int row = lane * R + r;
        x[r] = row < qlen ? q[row] : 0.0f;
    }

    int trow = qlen - 1;            // target cell
    int tpos = seg_lo + top_pos;
    int result = -1;
    int ck_limit = grp.n_ck;        // only checkpoints below this index may be used

    // The pass uses the score kernel's macro-steps (two columns per lane per step: lane l is on columns
    // 2(T-l), 2(T-l)+1 at macro-step T) so that its checkpoints can be resumed as they are.
    // Terminates: every retry restarts from a strictly earlier checkpoint (ck_limit shrinks), and the restart at
    // the segment's sentinel always resolves.
    for (;;) {
        // choose the restart: latest checkpoint k (< ck_limit) whose whole front lies at least
        // min_window columns before the target and after the segment's sentinel; else the sentinel.
        // Checkpoint k holds the state after macro-step T_k = 32*(k+1)*ck_every - 1: lane l's rows at
        // column 2(T_k - l) + 1.
        int k = -1;
        if (grp.ck_every > 0) {
            const long long lim = ((long long)tpos - a.min_window - 2) >> 1; // 2*T_k + 1 <= tpos - min_window - 1
            long long kk = lim >= 0 ? (lim + 1) / (32ll * grp.ck_every) - 1 : -1;
            if (kk >= ck_limit) kk = ck_limit - 1;
            if (kk >= 0) {
                const long long Tk = 32ll * (kk + 1) * grp.ck_every - 1;
                if (2 * (Tk - 31) >= seg_lo) k = (int)kk; // every front cell lies inside the segment
            }
        }
        float L[R];
        int S[R];
        float botA, botB, dprev;
        int sbotA, sbotB, sdprev;
        int T0; // first macro-step to execute
        int T = 0;
        if (k >= 0) {
            // Front cells carry an exit code instead of a start column: -1 - (4*row + off), the cell being
            // (row, 2(T - row/R) + 1 - off): off 0 = the lane's second column (L), off 1 = its first column
            // (bottom row only: botA), off 2 = the previous pair's second column (bottom row only: the next
            // lane's dprev).
            T = 32 * (k + 1) * grp.ck_every - 1;
            const float *c = a.ckpt + ((size_t)read * a.ck_per_read + grp.ck_prefix + k) * (size_t)a.ck_floats;
#pragma unroll
            for (int r = 0; r < R; r++) {
                L[r] = c[r * W + lane];
                S[r] = -1 - 4 * (lane * R + r);
            }
            dprev = c[R * W + lane];
            sdprev = -1 - (4 * (lane * R - 1) + 2);
            botA = c[(R + 1) * W + lane];
            sbotA = -1 - (4 * (lane * R + R - 1) + 1);
            T0 = T + 1;
        } else {
#pragma unroll
            for (int r = 0; r < R; r++) { L[r] = SF_INF; S[r] = 0; }
            dprev = (lane == 0 && !STD) ? 0.0f : SF_INF;
            sdprev = 0;
            botA = SF_INF;
            sbotA = 0;
            T0 = (seg_lo - 1) >> 1; // lane 0 starts on the pair that holds the sentinel
        }
        botB = L[R - 1];
        sbotB = S[R - 1];
        const int tl = trow / R, tr = trow % R;
        const int T_end = (tpos >> 1) + tl; // macro-step at which the target cell is produced
        const int tb = tpos & 1;            // target in the first (0) or second (1) column of the pair
        int sres = 0;
        // reference events: 32 pairs at a time, handed to the lanes by shuffle
        float yp0, yp1;
        {
            const long long c0 = 2ll * (T0 - 32 + lane_in_warp);
            yp0 = (c0 >= 0 && c0 < n_pos) ? __ldg(y + c0) : SF_INF;
            yp1 = (c0 + 1 >= 0 && c0 + 1 < n_pos) ? __ldg(y + c0 + 1) : SF_INF;
        }
        // one macro-step; CAP: also pick up the start pointers of row tr (only the step that produces the target)
        auto step = [&](const int s32, const int Tb, const float yc0, const float yc1, auto cap_tag) {
            constexpr bool CAP = decltype(cap_tag)::value;
            const int Tm = Tb + s32;
            const int colA = 2 * (Tm - lane);
            const int src = (s32 - lane) & 31;
            const float a0 = __shfl_sync(full, yc0, src), a1 = __shfl_sync(full, yc1, src);
            const float b0 = __shfl_sync(full, yp0, src), b1 = __shfl_sync(full, yp1, src);
            const float yA = s32 >= lane ? a0 : b0;
            const float yB = s32 >= lane ? a1 : b1;
            float upA = __shfl_up_sync(full, botA, 1, W);
            float upB = __shfl_up_sync(full, botB, 1, W);
            int supA = __shfl_up_sync(full, sbotA, 1, W);
            int supB = __shfl_up_sync(full, sbotB, 1, W);
            if (lane == 0) {
                upA = STD ? (yA == SF_INF ? 0.0f : SF_INF) : 0.0f;
                upB = STD ? (yB == SF_INF ? 0.0f : SF_INF) : 0.0f;
                supA = 0;
                supB = 0;
            }
            const float next_dprev = upB;
            const int next_sdprev = supB;
            float dgA = dprev, dgB = upA;
            int sdgA = sdprev, sdgB = supA;
            int capA = 0, capB = 0;
#pragma unroll
            for (int r = 0; r < R; r++) {
                // first column: left = L[r] (previous macro-step's second column)
                const float mA = fminf(fminf(upA, dgA), L[r]);
                int sA = sf_sel_eq(dgA, mA, sdgA, sf_sel_eq(L[r], mA, S[r], supA)); // diagonal, then left, then up
                if (r == 0 && lane == 0)
                    sA = colA - seg_lo; // start(0, j) = j
                const float va = fabsf(x[r] - yA) + mA;
                // second column: left = the value just computed
                const float mB = fminf(fminf(upB, dgB), va);
                int sB = sf_sel_eq(dgB, mB, sdgB, sf_sel_eq(va, mB, sA, supB));
                if (r == 0 && lane == 0)
                    sB = colA + 1 - seg_lo;
                const float vb = fabsf(x[r] - yB) + mB;
                dgA = L[r]; sdgA = S[r];
                dgB = va; sdgB = sA;
                L[r] = vb; S[r] = sB;
                upA = va; supA = sA;
                upB = vb; supB = sB;
                if (CAP) {
                    if (r == tr) { capA = sA; capB = sB; }
                }
            }
            dprev = next_dprev; sdprev = next_sdprev;
            botA = upA; sbotA = supA;
            botB = upB; sbotB = supB;
            if (CAP)
                sres = tb ? capB : capA;
        };
        const int n_blk = (T_end - T0) / 32 + 1;
        for (int blk = 0; blk < n_blk; blk++) {
            const int Tb = T0 + 32 * blk;
            const long long c0 = 2ll * (Tb + lane_in_warp);
            const float yc0 = (c0 >= 0 && c0 < n_pos) ? __ldg(y + c0) : SF_INF;
            const float yc1 = (c0 + 1 >= 0 && c0 + 1 < n_pos) ? __ldg(y + c0 + 1) : SF_INF;
            // the last block stops at the macro-step that produces the target
            const int n_plain = blk + 1 < n_blk ? 32 : T_end - Tb;
#pragma unroll 2
            for (int s32 = 0; s32 < n_plain; s32++)
                step(s32, Tb, yc0, yc1, sf_false());
            if (blk + 1 == n_blk)
                step(n_plain, Tb, yc0, yc1, sf_true());
            yp0 = yc0;
            yp1 = yc1;
        }
        sres = __shfl_sync(full, sres, tl);
        if (sres >= 0 || k < 0) {
            result = sres < 0 ? 0 : sres;
            break;
        }
        // the path left the window through the restart front: continue from that cell
        const int code = -1 - sres;
        trow = code >> 2;
        tpos = 2 * (T - trow / R) + 1 - (code & 3);
        ck_limit = k; // strictly earlier restart next time
    }
    return result;
}

// First attempt of the start-coordinate pass for TWO paired reads at once: lanes 0-15 hold the 16 x R rows of one
// read, lanes 16-31 those of another (the layout their checkpoints were taken in, sf_dtw_pair_kernel).  Every lane
// passes its own half's read / target / segment / group; `active` is false for a half with nothing to trace.
// The two halves run their own windows in lockstep (the warp iterates to the longer one; a half that is done
// keeps stepping on values nobody reads).  Returns the start column, or -2 when this half's path left its
// window through the restart front: the caller then runs the general sf_trace_start() for that read.
template <int R, bool STD>
__device__ __forceinline__ int sf_trace_first_dual(const sf_trace_args &a, const int read, const int lane_in_warp, const int qlen,
                                                   const int top_pos, const sf_seg &seg, const sf_group &grp, const bool active)
{
    constexpr int W = 16;
    const unsigned full = 0xffffffffu;
    const int lane = lane_in_warp & (W - 1);
    const int half_base = lane_in_warp & W; // 0 or 16
    const float *y = a.stream + grp.begin;
    const int n_pos = (int)(grp.end - grp.begin);
    const int seg_lo = (int)(seg.off - grp.begin);
    float x[R];
    const float *q = a.queries + (size_t)read * a.q_cap;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int row = lane * R + r;
        x[r] = row < qlen ? q[row] : 0.0f;
    }
    const int trow = qlen - 1;
    const int tpos = seg_lo + top_pos;
    // restart: as in sf_trace_start()
    int k = -1;
    if (grp.ck_every > 0) {
        const long long lim = ((long long)tpos - a.min_window - 2) >> 1;
        long long kk = lim >= 0 ? (lim + 1) / (32ll * grp.ck_every) - 1 : -1;
        if (kk >= grp.n_ck) kk = grp.n_ck - 1;
        if (kk >= 0) {
            const long long Tk = 32ll * (kk + 1) * grp.ck_every - 1;
            if (2 * (Tk - 31) >= seg_lo) k = (int)kk; // same rule as sf_trace_start()
        }
    }
    float L[R];
    int S[R];
    float botA, botB, dprev;
    int sbotA, sbotB, sdprev;
    int T0;
    if (k >= 0) {
        const int T = 32 * (k + 1) * grp.ck_every - 1;
        const float *c = a.ckpt + ((size_t)read * a.ck_per_read + grp.ck_prefix + k) * (size_t)a.ck_floats;
#pragma unroll
        for (int r = 0; r < R; r++) {
            L[r] = c[r * W + lane];
            S[r] = -1 - 4 * (lane * R + r);
        }
        dprev = c[R * W + lane];
        sdprev = -1 - (4 * (lane * R - 1) + 2);
        botA = c[(R + 1) * W + lane];
        sbotA = -1 - (4 * (lane * R + R - 1) + 1);
        T0 = T + 1;
    } else {
#pragma unroll
        for (int r = 0; r < R; r++) { L[r] = SF_INF; S[r] = 0; }
        dprev = (lane == 0 && !STD) ? 0.0f : SF_INF;
        sdprev = 0;
        botA = SF_INF;
        sbotA = 0;
        T0 = (seg_lo - 1) >> 1;
    }
    botB = L[R - 1];
    sbotB = S[R - 1];
    const int tl = trow / R, tr = trow % R;
    const int T_end = (tpos >> 1) + tl;
    const int tb = tpos & 1;
    // macro-steps of this half (the last one produces the target) and of the other one
    const int e_mine = active ? T_end - T0 : -1;
    const int e_other = __shfl_xor_sync(full, e_mine, W);
    const int n_steps = max(e_mine, e_other) + 1;
    int sres = 0;
    // reference events: 16 pairs at a time per half, handed to the lanes by shuffle
    float yp0, yp1, yc0 = SF_INF, yc1 = SF_INF;
    {
        const long long c0 = 2ll * (T0 - W + lane);
        yp0 = (c0 >= 0 && c0 < n_pos) ? __ldg(y + c0) : SF_INF;
        yp1 = (c0 + 1 >= 0 && c0 + 1 < n_pos) ? __ldg(y + c0 + 1) : SF_INF;
    }
    auto step = [&](const int i, auto cap_tag) {
        constexpr bool CAP = decltype(cap_tag)::value;
        const int s16 = i & (W - 1);
        const int Tm = T0 + i;
        const int colA = 2 * (Tm - lane);
        const int src = half_base + ((s16 - lane) & (W - 1));
        const float a0 = __shfl_sync(full, yc0, src), a1 = __shfl_sync(full, yc1, src);
        const float b0 = __shfl_sync(full, yp0, src), b1 = __shfl_sync(full, yp1, src);
        const float yA = s16 >= lane ? a0 : b0;
        const float yB = s16 >= lane ? a1 : b1;
        float upA = __shfl_up_sync(full, botA, 1, W);
        float upB = __shfl_up_sync(full, botB, 1, W);
        int supA = __shfl_up_sync(full, sbotA, 1, W);
        int supB = __shfl_up_sync(full, sbotB, 1, W);
        if (lane == 0) {
            upA = STD ? (yA == SF_INF ? 0.0f : SF_INF) : 0.0f;
            upB = STD ? (yB == SF_INF ? 0.0f : SF_INF) : 0.0f;
            supA = 0;
            supB = 0;
        }
        const float next_dprev = upB;
        const int next_sdprev = supB;
        float dgA = dprev, dgB = upA;
        int sdgA = sdprev, sdgB = supA;
        int capA = 0, capB = 0;
#pragma unroll
        for (int r = 0; r < R; r++) {
            const float mA = fminf(fminf(upA, dgA), L[r]);
            int sA = sf_sel_eq(dgA, mA, sdgA, sf_sel_eq(L[r], mA, S[r], supA)); // diagonal, then left, then up
            if (r == 0 && lane == 0)
                sA = colA - seg_lo; // start(0, j) = j
            const float va = fabsf(x[r] - yA) + mA;
            const float mB = fminf(fminf(upB, dgB), va);
            int sB = sf_sel_eq(dgB, mB, sdgB, sf_sel_eq(va, mB, sA, supB));
            if (r == 0 && lane == 0)
                sB = colA + 1 - seg_lo;
            const float vb = fabsf(x[r] - yB) + mB;
            dgA = L[r]; sdgA = S[r];
            dgB = va; sdgB = sA;
            L[r] = vb; S[r] = sB;
            upA = va; supA = sA;
            upB = vb; supB = sB;
            if (CAP) {
                if (r == tr) { capA = sA; capB = sB; }
            }
        }
        dprev = next_dprev; sdprev = next_sdprev;
        botA = upA; sbotA = supA;
        botB = upB; sbotB = supB;
        if (CAP) {
            if (i == e_mine)
                sres = tb ? capB : capA;
        }
    };
    for (int i = 0; i < n_steps; i++) {
        if ((i & (W - 1)) == 0) { // next 16 pairs of this half's reference events
            if (i > 0) { yp0 = yc0; yp1 = yc1; }
            const long long c0 = 2ll * (T0 + i + lane);
            yc0 = (c0 >= 0 && c0 < n_pos) ? __ldg(y + c0) : SF_INF;
            yc1 = (c0 + 1 >= 0 && c0 + 1 < n_pos) ? __ldg(y + c0 + 1) : SF_INF;
        }
        if (i == e_mine || i == e_other) // warp-uniform: both values are known to every lane
            step(i, sf_true());
        else
            step(i, sf_false());
    }
    sres = __shfl_sync(full, sres, half_base + tl);
    if (!active)
        return -1;
    if (sres >= 0 || k < 0)
        return sres < 0 ? 0 : sres;
    return -2;
}

// The pass is latency bound (one dependent chain of 2R+1 cells per macro-step), so resident warps matter more than
// registers: up to 16 rows per lane the kernel is held to 128 registers (4 blocks of 4 warps per SM).
__host__ __device__ constexpr int sf_trace_min_blocks(int R, int R2) { return (R > R2 ? R : R2) <= 16 ? 4 : 1; }

// R2 > 0: reads flagged with status bit 5 were aligned by sf_dtw_pair_kernel<R2> (half-warp checkpoints)
template <int R, bool STD, int R2>
__global__ void __launch_bounds__(128, sf_trace_min_blocks(R, R2)) sf_trace_kernel(const sf_trace_args a)
{
    const int lane = threadIdx.x & 31;
    const int read = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const unsigned full = 0xffffffffu;
    if (read >= a.n_reads)
        return;
    const int qlen = a.info[read].qlen;
    if (R2 > 0 && (a.info[read].status & 32))
        return; // a paired read: sf_trace_pair_kernel
    sf_hit hit;
    hit.score = SF_INF; hit.score2 = SF_INF; hit.rid = -1; hit.strand = 0;
    hit.pos_st = -1; hit.pos_end = -1; hit.seg = -1; hit.pad = 0;
    if (qlen <= 0) {
        if (lane == 0) a.hits[read] = hit;
        return;
    }

    // ---- merge the per-group results of this read ----
    sf_top top;
    top.s1 = SF_INF; top.s2 = SF_INF; top.seg = -1; top.chunk = 0; top.pos = -1;
    for (int t = lane; t < a.n_pieces; t += 32)
        sf_merge_task(top, a, a.res + (size_t)read * a.n_pieces, t);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sf_top b;
        b.s1 = __shfl_xor_sync(full, top.s1, o);
        b.s2 = __shfl_xor_sync(full, top.s2, o);
        b.seg = __shfl_xor_sync(full, top.seg, o);
        b.chunk = __shfl_xor_sync(full, top.chunk, o);
        b.pos = __shfl_xor_sync(full, top.pos, o);
        sf_top_merge(top, b);
    }
    hit.score = top.s1; hit.score2 = top.s2; hit.seg = top.seg;
    if (top.seg < 0 || top.pos < 0) {
        if (top.seg >= 0) {
            const sf_seg sg = a.segs[top.seg];
            hit.rid = sg.rid; hit.strand = sg.strand; hit.pos_end = top.pos;
        }
        if (lane == 0) a.hits[read] = hit;
        return;
    }
    const sf_seg seg = a.segs[top.seg];
    hit.rid = seg.rid; hit.strand = seg.strand; hit.pos_end = top.pos;

    // ---- start-coordinate pass ----
    const int gid = a.seg_group[top.seg];
    const sf_group grp = a.groups[gid];

    hit.pos_st = sf_trace_start<R, STD, 32>(a, read, lane, qlen, top.pos, seg, grp);
    if (lane == 0) a.hits[read] = hit;
}

// Merge + start coordinate of the paired reads, two per warp (reads list_full[2w], list_full[2w+1]; R rows per
// lane, 16 lanes per read).
template <int R, bool STD>
__global__ void __launch_bounds__(128, sf_trace_min_blocks(R, 0)) sf_trace_pair_kernel(const sf_trace_args a)
{
    const int lane = threadIdx.x & 31;
    const int ll = lane & 15, half = lane >> 4;
    const int unit = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const unsigned full = 0xffffffffu;
    const int n_full = *a.n_full;
    if (2 * unit >= n_full)
        return;
    const bool present = 2 * unit + half < n_full; // an odd list leaves the last upper half empty
    const int read = a.list_full[present ? 2 * unit + half : 2 * unit];
    const int qlen = a.info[read].qlen;
    sf_hit hit;
    hit.score = SF_INF; hit.score2 = SF_INF; hit.rid = -1; hit.strand = 0;
    hit.pos_st = -1; hit.pos_end = -1; hit.seg = -1; hit.pad = 0;

    // ---- merge the per-group results, 16 lanes per read ----
    sf_top top;
    top.s1 = SF_INF; top.s2 = SF_INF; top.seg = -1; top.chunk = 0; top.pos = -1;
    for (int t = ll; t < a.n_pieces; t += 16)
        sf_merge_task(top, a, a.res + (size_t)read * a.n_pieces, t);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        sf_top b;
        b.s1 = __shfl_xor_sync(full, top.s1, o);
        b.s2 = __shfl_xor_sync(full, top.s2, o);
        b.seg = __shfl_xor_sync(full, top.seg, o);
        b.chunk = __shfl_xor_sync(full, top.chunk, o);
        b.pos = __shfl_xor_sync(full, top.pos, o);
        sf_top_merge(top, b);
    }
    hit.score = top.s1; hit.score2 = top.s2; hit.seg = top.seg;
    const bool traced = present && top.seg >= 0 && top.pos >= 0;
    sf_seg seg;
    seg.off = 0; seg.rlen = 0; seg.rid = -1; seg.strand = 0;
    sf_group grp = a.groups[0];
    if (top.seg >= 0) {
        seg = a.segs[top.seg];
        hit.rid = seg.rid; hit.strand = seg.strand; hit.pos_end = top.pos;
        grp = a.groups[a.seg_group[top.seg]];
    }
    int result = sf_trace_first_dual<R, STD>(a, read, lane, qlen, traced ? top.pos : 0, seg, grp, traced);
    // a path that left its window: the general pass, one read at a time (the warp mirrors the read in both halves)
#pragma unroll 1
    for (int h = 0; h < 2; h++) {
        if (__shfl_sync(full, result, 16 * h) != -2)
            continue;
        const int read_h = __shfl_sync(full, read, 16 * h);
        const int pos_h = __shfl_sync(full, top.pos, 16 * h);
        const int seg_h = __shfl_sync(full, top.seg, 16 * h);
        const sf_seg sg = a.segs[seg_h];
        const sf_group gp = a.groups[a.seg_group[seg_h]];
        const int r = sf_trace_start<R, STD, 16>(a, read_h, lane, qlen, pos_h, sg, gp);
        if (half == h)
            result = r;
    }
    if (traced)
        hit.pos_st = result;
    if (ll == 0 && present) a.hits[read] = hit;
}
