// ubench_alu2.cu -- second issue-rate microbenchmark: packed fp32 (FADD2) and integer-min variants of
// the sDTW cell.  Reports cycles per cell per SMSP (lower is better) at several warp counts.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o ubench_alu2 ubench_alu2.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define P 4        // cell pairs per thread per "step" (= R/2 for R = 8)
#define ITERS 4096

__device__ __forceinline__ float imin3(float a, float b, float c)
{
    return __int_as_float(min(min(__float_as_int(a), __float_as_int(b)), __float_as_int(c)));
}

// MODE 0: scalar cell x2P          (FADD, FMNMX3, FADD|.|)
// MODE 1: FADD2 t, 2 FMNMX3, 2 FADD|.|
// MODE 2: FADD2 t, 2 FMNMX3, FADD2 |t|+m
// MODE 3: FADD2 t, 2 integer min3, FADD2 |t|+m
// MODE 4: FADD2 only (2 per pair)
// MODE 5: integer min3 only (2 per pair)
// MODE 6: FADD2 t, FMNMX3 + imin3 (one of each), FADD2
template <int MODE> __global__ void k(float *out, long long *cycles, float seed)
{
    float2 L[P], x[P];
#pragma unroll
    for (int i = 0; i < P; i++) {
        L[i] = make_float2(seed + threadIdx.x * 0.001f + i, seed * 2 + i);
        x[i] = make_float2(seed * 0.5f + i, seed * 0.25f - i);
    }
    float2 yy = make_float2(-seed * 0.25f, -seed * 0.125f);
    float upA = seed, upB = seed * 3, dA = seed * 5, dB = seed * 7;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        float dgA = dA, dgB = dB;
        float uA = upA, uB = upB;
#pragma unroll
        for (int p = 0; p < P; p++) {
            if (MODE == 0) {
                const float tA = x[p].x + yy.x, tB = x[p].y + yy.y;
                const float mA = fminf(fminf(uA, dgA), L[p].x), mB = fminf(fminf(uB, dgB), L[p].y);
                dgA = L[p].x; dgB = L[p].y;
                L[p].x = fabsf(tA) + mA; L[p].y = fabsf(tB) + mB;
                uA = L[p].x; uB = L[p].y;
            } else if (MODE == 1) {
                const float2 t = __fadd2_rn(x[p], yy);
                const float mA = fminf(fminf(uA, dgA), L[p].x), mB = fminf(fminf(uB, dgB), L[p].y);
                dgA = L[p].x; dgB = L[p].y;
                L[p].x = fabsf(t.x) + mA; L[p].y = fabsf(t.y) + mB;
                uA = L[p].x; uB = L[p].y;
            } else if (MODE == 2 || MODE == 3 || MODE == 6) {
                const float2 t = __fadd2_rn(x[p], yy);
                float mA, mB;
                if (MODE == 2) { mA = fminf(fminf(uA, dgA), L[p].x); mB = fminf(fminf(uB, dgB), L[p].y); }
                else if (MODE == 3) { mA = imin3(uA, dgA, L[p].x); mB = imin3(uB, dgB, L[p].y); }
                else { mA = fminf(fminf(uA, dgA), L[p].x); mB = imin3(uB, dgB, L[p].y); }
                dgA = L[p].x; dgB = L[p].y;
                L[p] = __fadd2_rn(make_float2(fabsf(t.x), fabsf(t.y)), make_float2(mA, mB));
                uA = L[p].x; uB = L[p].y;
            } else if (MODE == 4) {
                const float2 t = __fadd2_rn(x[p], yy);
                L[p] = __fadd2_rn(make_float2(fabsf(t.x), fabsf(t.y)), L[p]);
            } else if (MODE == 5) {
                L[p].x = imin3(uA, dgA, L[p].x); L[p].y = imin3(uB, dgB, L[p].y);
                dgA = uA; dgB = uB; uA = L[p].x + 0.0f * 0; uB = L[p].y;
            }
        }
        dA = upA; dB = upB;
        upA = L[P - 1].y; upB = L[P - 1].x;
        yy.x += 1e-7f;
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < P; i++) s += L[i].x + L[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + dA + dB;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char *name, int warps_per_smsp)
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int threads = warps_per_smsp * 4 * 32;
    const int blocks = p.multiProcessorCount;
    float *out;
    long long *cyc;
    cudaMalloc(&out, sizeof(float) * blocks * threads);
    cudaMalloc(&cyc, sizeof(long long) * blocks);
    for (int r = 0; r < 2; r++) {
        k<MODE><<<blocks, threads>>>(out, cyc, 1.0f);
        cudaDeviceSynchronize();
    }
    long long h[1024];
    cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; i++) avg += (double)h[i];
    avg /= blocks;
    const double cells = (double)ITERS * 2 * P * warps_per_smsp; // warp-cells per SMSP
    printf("%-44s warps/SMSP=%2d  cycles/warp-cell/SMSP=%.3f\n", name, warps_per_smsp, avg / cells);
    cudaFree(out);
    cudaFree(cyc);
}

int main()
{
    for (int w : {2, 4, 8}) {
        run<0>("0 scalar FADD,FMNMX3,FADD", w);
        run<1>("1 FADD2 t; 2 FMNMX3; 2 FADD", w);
        run<2>("2 FADD2 t; 2 FMNMX3; FADD2", w);
        run<3>("3 FADD2 t; 2 int-min3; FADD2", w);
        run<6>("6 FADD2 t; FMNMX3+int-min3; FADD2", w);
        run<4>("4 2 FADD2 only", w);
        run<5>("5 2 int-min3 only", w);
    }
    return 0;
}
