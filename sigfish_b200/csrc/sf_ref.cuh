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
// accumulated in the reference's order by one lane from tiles the warp stages in shared memory), scale (parallel).
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

// One warp per segment; serial fp32 accumulation order of genref.c:28-41.  fp32 addition is not associative, so
// the two sums are chains of dependent FADDs and cannot be split; what can be done is to make the chain cost one
// FADD latency per element and nothing else: the warp stages tiles of the array (pass 2: of (x - mean)^2) in
// shared memory with coalesced loads, prefetching the next tile into registers, and lane 0 runs the chain over the
// staged tile with 16-byte shared loads that do not depend on the running sum.  ~5 cycles per element instead of ~40
// (a 100 Mb contig: 2 x 1e8 dependent adds per strand, well under a second).
#define SF_REF_TILE 1024
#define SF_REF_STAT_WARPS 4
__global__ void __launch_bounds__(32 * SF_REF_STAT_WARPS) sf_ref_stats_kernel(const sf_ref_args a)
{
    __shared__ __align__(16) float tile_all[SF_REF_STAT_WARPS][SF_REF_TILE];
    const int lane = threadIdx.x & 31;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    float *tile = tile_all[threadIdx.x >> 5];
    constexpr int PER = SF_REF_TILE / 32;
    for (int s = wid; s < a.n_seg; s += nw) {
        const sf_seg sg = a.segs[s];
        const float *v = a.stream + sg.off;
        const int n = sg.rlen;
        const float cnt = (float)(unsigned long long)n;
        float mean = 0.0f, acc = 0.0f;
#pragma unroll 1
        for (int pass = 0; pass < 2; pass++) {
            acc = 0.0f;
            float nxt[PER];
#pragma unroll
            for (int k = 0; k < PER; k++) {
                const int i = k * 32 + lane;
                nxt[k] = i < n ? v[i] : 0.0f;
            }
            for (int b = 0; b < n; b += SF_REF_TILE) {
                __syncwarp();
#pragma unroll
                for (int k = 0; k < PER; k++) {
                    float x = nxt[k];
                    if (pass == 1) {
                        const float d = __fsub_rn(x, mean);
                        x = __fmul_rn(d, d);
                    }
                    tile[k * 32 + lane] = x;
                }
                __syncwarp();
#pragma unroll
                for (int k = 0; k < PER; k++) { // next tile: in flight while lane 0 runs the chain
                    const int i = b + SF_REF_TILE + k * 32 + lane;
                    nxt[k] = i < n ? v[i] : 0.0f;
                }
                if (lane == 0) {
                    const int lim = min(SF_REF_TILE, n - b);
                    int j = 0;
                    for (; j + 4 <= lim; j += 4) {
                        const float4 t = *reinterpret_cast<const float4 *>(tile + j);
                        acc = __fadd_rn(acc, t.x);
                        acc = __fadd_rn(acc, t.y);
                        acc = __fadd_rn(acc, t.z);
                        acc = __fadd_rn(acc, t.w);
                    }
                    for (; j < lim; j++)
                        acc = __fadd_rn(acc, tile[j]);
                }
            }
            acc = __shfl_sync(0xffffffffu, acc, 0);
            if (pass == 0)
                mean = __fdiv_rn(acc, cnt);
        }
        const float var = __fdiv_rn(acc, cnt);
        if (lane == 0)
            a.stats[s] = make_float2(mean, __fsqrt_rn(var));
        __syncwarp();
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
