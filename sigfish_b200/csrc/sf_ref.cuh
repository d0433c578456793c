// sf_ref.cuh -- kernel #2: k-mer-model reference synthesis, built once and kept in HBM.
//
// Replaces (reference, paths relative to /root/reference):
//   src/ref.h:13-41     get_rank / get_kmer_rank
//   src/ref.h:45-76     complement / reverse_complement
//   src/genref.c:157-198 forward (+ reverse-complement for DNA) level_mean lookup, RNA window
//                        variants (--from-end, --invert)
//   src/genref.c:23-47  per-array z-score (fp32 running sums in array order)
//
// Three launches: fill (parallel over columns), stats (one warp per segment; the fp32 sums are
// accumulated in the reference's order -- every lane replays the same serial chain from values
// that were loaded coalesced and exchanged by shuffle), scale (parallel).
#pragma once
#include <cuda_runtime.h>
#include "sf_types.cuh"

struct sf_refseg {      // how to synthesise one segment
    int64_t base_off;   // offset of the contig's first base in `bases`
    int32_t seq_len;    // contig length l
    int32_t first;      // index of the first k-mer used on the forward string (RNA windows)
    int32_t mode;       // 0 forward, 1 reverse complement, 2 forward written back to front (--invert)
    int32_t pad;
};

struct sf_ref_args {
    const uint8_t *bases;
    const sf_refseg *rseg;
    const sf_seg *segs;
    int32_t n_seg;
    const float *level_mean;
    int32_t k;
    float *stream;
    float2 *stats;      // per segment (mean, stdv)
};

__device__ __forceinline__ uint32_t sf_base_code(uint8_t b)
{
    switch (b) {
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 0; // A, a and everything else (ref.h:23-25)
    }
}

// code of the complement: A<->T, C<->G, unknown -> 'T' (ref.h:62-65), i.e. rank 3
__device__ __forceinline__ uint32_t sf_comp_code(uint8_t b)
{
    switch (b) {
    case 'A': case 'a': return 3;
    case 'C': case 'c': return 2;
    case 'G': case 'g': return 1;
    case 'T': case 't': return 0;
    default: return 3;
    }
}

__global__ void sf_ref_fill_kernel(const sf_ref_args a)
{
    for (int s = blockIdx.y; s < a.n_seg; s += gridDim.y) {
        const sf_refseg rs = a.rseg[s];
        const sf_seg sg = a.segs[s];
        const uint8_t *seq = a.bases + rs.base_off;
        float *out = a.stream + sg.off;
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < sg.rlen; j += gridDim.x * blockDim.x) {
            uint32_t rank = 0;
            if (rs.mode == 1) {
                // k-mer j of the reverse complement: comp(seq[l-1-j]), comp(seq[l-2-j]), ...
                const int top = rs.seq_len - 1 - j;
                for (int i = 0; i < a.k; i++)
                    rank = (rank << 2) | sf_comp_code(seq[top - i]);
                out[j] = a.level_mean[rank];
            } else {
                const int st = rs.first + j;
                for (int i = 0; i < a.k; i++)
                    rank = (rank << 2) | sf_base_code(seq[st + i]);
                out[rs.mode == 2 ? sg.rlen - 1 - j : j] = a.level_mean[rank];
            }
        }
    }
}

// one warp per segment; serial fp32 accumulation order of genref.c:28-41
__global__ void sf_ref_stats_kernel(const sf_ref_args a)
{
    const int lane = threadIdx.x & 31;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    const unsigned full = 0xffffffffu;
    for (int s = wid; s < a.n_seg; s += nw) {
        const sf_seg sg = a.segs[s];
        const float *v = a.stream + sg.off;
        const int n = sg.rlen;
        float sum = 0.0f;
        for (int b = 0; b < n; b += 32) {
            const float mine = (b + lane < n) ? v[b + lane] : 0.0f;
            const int lim = min(32, n - b);
            for (int j = 0; j < lim; j++)
                sum = __fadd_rn(sum, __shfl_sync(full, mine, j));
        }
        const float cnt = (float)(unsigned long long)n;
        const float mean = __fdiv_rn(sum, cnt);
        float var = 0.0f;
        for (int b = 0; b < n; b += 32) {
            const float mine = (b + lane < n) ? v[b + lane] : 0.0f;
            const float d = __fsub_rn(mine, mean);
            const float d2 = __fmul_rn(d, d);
            const int lim = min(32, n - b);
            for (int j = 0; j < lim; j++)
                var = __fadd_rn(var, __shfl_sync(full, d2, j));
        }
        var = __fdiv_rn(var, cnt);
        if (lane == 0)
            a.stats[s] = make_float2(mean, __fsqrt_rn(var));
    }
}

__global__ void sf_ref_scale_kernel(const sf_ref_args a)
{
    for (int s = blockIdx.y; s < a.n_seg; s += gridDim.y) {
        const sf_seg sg = a.segs[s];
        const float2 st = a.stats[s];
        float *v = a.stream + sg.off;
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < sg.rlen; j += gridDim.x * blockDim.x)
            v[j] = __fdiv_rn(__fsub_rn(v[j], st.x), st.y);
    }
}

__global__ void sf_fill_inf_kernel(float *p, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = SF_INF;
}
