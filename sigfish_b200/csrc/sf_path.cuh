// sf_path.cuh -- --sam only: the full warping path of each read's winning hit.
//
// Replaces (reference, paths relative to /root/reference):
//   src/cdtw.c:98-167   path()             -- backtrack through the stored cost matrix
//   src/cdtw.c:192-227  subsequence_path() -- drop the leading row-0 run
// The reference keeps the whole qlen x rlen matrix; here only the columns the path can visit are
// recomputed: [pos_st, pos_end] of the winning segment for subsequence DTW (pos_st is already known
// from sf_trace.cuh), [0, pos_end] for standard DTW (whose row 0 is cumulative).  For subsequence DTW
// the window is entered through a +INF column: every cell ON the winner's path keeps its exact value
// (its optimal path lies inside the window) and every other cell can only grow, so each equality test
// of the backtrack (diagonal, then left, then up -- cdtw.c:134-146) decides as on the full matrix.
// Forward pass: the same skewed warp wavefront as the score kernel, each lane packing 2 bits per cell
// (0 diagonal, 1 left, 2 up) into one 64-bit word per column.  Backward pass: lane 0 walks the bits
// from (qlen-1, pos_end) up to row 0 and emits the moves; the host rebuilds (px, py) and the
// reference-to-event map (sigfish.c:530-571).
#pragma once
#include <cuda_runtime.h>
#include "sf_types.cuh"

struct sf_path_args {
    const float *stream;
    const sf_seg *segs;
    const float *queries;
    const sf_readinfo *info;
    int32_t q_cap;
    const sf_hit *hits;
    int32_t n_reads;
    const int64_t *dir_off;        // [n_reads] first column slot of the read in dirs, -1: nothing to do
    unsigned long long *dirs;      // [columns][32]
    const int64_t *move_off;       // [n_reads + 1] offsets into moves
    uint8_t *moves;                // backward order: from the end cell towards row 0
    int32_t *n_moves;              // [n_reads]; -1: no path
    int32_t *start_col;            // [n_reads] column at which row 0 was reached (must equal pos_st)
};

template <int R, bool STD>
__global__ void __launch_bounds__(128) sf_path_kernel(const sf_path_args a)
{
    static_assert(R <= 32, "2 bits x R rows must fit one 64-bit word");
    const int lane = threadIdx.x & 31;
    const int read = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const unsigned full = 0xffffffffu;
    if (read >= a.n_reads)
        return;
    const sf_hit hit = a.hits[read];
    const int qlen = a.info[read].qlen;
    const long long doff = a.dir_off[read];
    if (qlen <= 0 || hit.seg < 0 || hit.pos_st < 0 || hit.pos_end < hit.pos_st || doff < 0) {
        if (lane == 0) {
            a.n_moves[read] = -1;
            a.start_col[read] = -1;
        }
        return;
    }
    const sf_seg seg = a.segs[hit.seg];
    const int c0 = STD ? 0 : hit.pos_st;
    const int width = hit.pos_end - c0 + 1;
    const float *y = a.stream + seg.off + c0;
    unsigned long long *dirs = a.dirs + (size_t)doff * 32;

    float x[R], L[R];
    const float *q = a.queries + (size_t)read * a.q_cap;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int row = lane * R + r;
        x[r] = row < qlen ? q[row] : 0.0f;
        L[r] = SF_INF;
    }
    float bot = SF_INF;
    // virtual row -1: all zeros for subsequence DTW; for standard DTW only the corner (-1,-1) is 0
    float dprev = lane == 0 ? 0.0f : SF_INF;
    const int lq = (qlen - 1) / R;
    const int t_last = width - 1 + lq;
    for (int t = 0; t <= t_last; t++) {
        const int pos = t - lane;
        const bool live = pos >= 0 && pos < width;
        const float yy = live ? __ldg(y + pos) : SF_INF;
        float up = __shfl_up_sync(full, bot, 1);
        if (lane == 0)
            up = STD ? SF_INF : 0.0f;
        const float unext = up;
        float dg = dprev;
        unsigned long long code = 0;
#pragma unroll
        for (int r = 0; r < R; r++) {
            const float m = fminf(fminf(up, dg), L[r]);
            const unsigned d = (dg == m) ? 0u : ((L[r] == m) ? 1u : 2u);
            code |= (unsigned long long)d << (2 * r);
            const float nv = fabsf(x[r] - yy) + m;
            dg = L[r];
            L[r] = nv;
            up = nv;
        }
        dprev = unext;
        bot = L[R - 1];
        if (live)
            dirs[(size_t)pos * 32 + lane] = code;
    }
    __threadfence_block();
    __syncwarp();
    if (lane == 0) {
        uint8_t *mv = a.moves + a.move_off[read];
        const long long cap = a.move_off[read + 1] - a.move_off[read];
        int i = qlen - 1, j = width - 1;
        long long n = 0;
        while (i > 0 && n < cap) {
            unsigned d;
            if (j == 0) {
                d = 2u; // cdtw.c:130-131
            } else {
                const unsigned long long code = dirs[(size_t)j * 32 + i / R];
                d = (unsigned)(code >> (2 * (i % R))) & 3u;
            }
            mv[n++] = (uint8_t)d;
            if (d == 0u) { i--; j--; }
            else if (d == 1u) { j--; }
            else { i--; }
        }
        a.n_moves[read] = i == 0 ? (int32_t)n : -1;
        a.start_col[read] = c0 + j;
    }
}
