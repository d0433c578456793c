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
// that front cell as the new target.
#pragma once
#include <cuda_runtime.h>
#include "sf_types.cuh"

struct sf_trace_args {
    const float *stream;
    const sf_seg *segs;
    const sf_group *groups;
    const int32_t *seg_group; // segment -> group id
    int32_t n_groups;
    int32_t n_reads;
    const float *queries;
    const sf_readinfo *info;
    int32_t q_cap;
    const sf_taskres *res;
    const float *ckpt;
    int64_t ck_per_read;
    sf_hit *hits;
    int32_t min_window;       // restart at least this many columns before the target
};

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

template <int R, bool STD>
__global__ void __launch_bounds__(128) sf_trace_kernel(const sf_trace_args a)
{
    const int lane = threadIdx.x & 31;
    const int read = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const unsigned full = 0xffffffffu;
    if (read >= a.n_reads)
        return;
    const int qlen = a.info[read].qlen;
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
    for (int g = lane; g < a.n_groups; g += 32) {
        const sf_taskres tr = a.res[(size_t)read * a.n_groups + g];
        sf_top b; b.s1 = tr.s1; b.s2 = tr.s2; b.seg = tr.seg; b.chunk = tr.chunk; b.pos = tr.pos;
        sf_top_merge(top, b);
    }
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
    int tpos = seg_lo + top.pos;
    int result = -1;
    int ck_limit = grp.n_ck;        // only checkpoints below this index may be used

    for (int attempt = 0; attempt < 64; attempt++) {
        // choose the restart: latest checkpoint k (< ck_limit) whose whole front lies at least
        // min_window columns before the target and after the segment's sentinel; else the sentinel
        int k = -1;
        if (grp.ck_every > 0) {
            // checkpoint k holds the state after step T_k = 32*(k+1)*ck_every - 1
            const long long lim = (long long)tpos - a.min_window - 1;
            long long kk = (lim + 1) / (32ll * grp.ck_every) - 1;
            if (kk >= ck_limit) kk = ck_limit - 1;
            if (kk >= 0) {
                const long long Tk = 32ll * (kk + 1) * grp.ck_every - 1;
                if (Tk - 31 > seg_lo - 1) k = (int)kk;
            }
        }
        float L[R];
        int S[R];
        float bot, dprev;
        int sbot, sdprev;
        int t0; // first step to execute
        int T = 0;
        if (k >= 0) {
            T = 32 * (k + 1) * grp.ck_every - 1;
            const float *c = a.ckpt + ((size_t)read * a.ck_per_read + grp.ck_prefix + k) * (size_t)((R + 1) * 32);
#pragma unroll
            for (int r = 0; r < R; r++) {
                L[r] = c[r * 32 + lane];
                S[r] = -1 - 2 * (lane * R + r);          // front cell (row, T - lane)
            }
            dprev = c[R * 32 + lane];
            sdprev = -1 - (2 * (lane * R - 1) + 1);      // front cell (lane*R - 1, T - lane)
            t0 = T + 1;
        } else {
#pragma unroll
            for (int r = 0; r < R; r++) { L[r] = SF_INF; S[r] = 0; }
            dprev = (lane == 0 && !STD) ? 0.0f : SF_INF;
            sdprev = 0;
            t0 = seg_lo - 1; // lane 0 starts on the sentinel
        }
        bot = L[R - 1];
        sbot = S[R - 1];
        const int tl = trow / R, tr = trow % R;
        const int t_end = tpos + tl; // step at which the target cell is produced
        int sres = 0;
        // reference events are fetched 32 steps at a time (one coalesced load per lane) and handed to the
        // lanes by shuffle: lane l needs position t - l, which sits in this block's or the previous block's load
        float yprev;
        {
            const int pp = t0 - 32 + lane;
            yprev = (pp >= 0 && pp < n_pos) ? __ldg(y + pp) : SF_INF;
        }
        const int n_blk = (t_end - t0) / 32 + 1;
        for (int blk = 0; blk < n_blk; blk++) {
            const int tb = t0 + 32 * blk;
            const int pc = tb + lane;
            const float ycur = (pc >= 0 && pc < n_pos) ? __ldg(y + pc) : SF_INF;
#pragma unroll 4
            for (int s32 = 0; s32 < 32; s32++) {
                const int t = tb + s32;
                const int pos = t - lane;
                const int src = (s32 - lane) & 31;
                const float ya = __shfl_sync(full, ycur, src);
                const float yb = __shfl_sync(full, yprev, src);
                const float yy = s32 >= lane ? ya : yb;
                float up = __shfl_up_sync(full, bot, 1);
                int sup = __shfl_up_sync(full, sbot, 1);
                if (lane == 0) {
                    up = STD ? (yy == SF_INF ? 0.0f : SF_INF) : 0.0f;
                    sup = 0;
                }
                const float unext = up;
                const int sunext = sup;
                float dg = dprev;
                int sdg = sdprev;
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const float m = fminf(fminf(up, dg), L[r]);
                    int s = (dg == m) ? sdg : ((L[r] == m) ? S[r] : sup);
                    if (lane == 0 && r == 0)
                        s = pos - seg_lo; // start(0, j) = j
                    const float nv = fabsf(x[r] - yy) + m;
                    dg = L[r]; sdg = S[r];
                    L[r] = nv; S[r] = s;
                    up = nv; sup = s;
                }
                dprev = unext; sdprev = sunext;
                bot = L[R - 1]; sbot = S[R - 1];
                if (t == t_end) {
#pragma unroll
                    for (int r = 0; r < R; r++)
                        if (r == tr) sres = S[r];
                }
            }
            yprev = ycur;
        }
        sres = __shfl_sync(full, sres, tl);
        if (sres >= 0 || k < 0) {
            result = sres < 0 ? 0 : sres;
            break;
        }
        // the path left the window through the restart front: continue from that cell
        const int code = -1 - sres;
        trow = code >> 1;
        tpos = T - trow / R - (code & 1);
        ck_limit = k; // strictly earlier restart next time
    }
    hit.pos_st = result;
    if (lane == 0) a.hits[read] = hit;
}
