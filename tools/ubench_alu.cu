// ubench_alu.cu -- issue-rate microbenchmark for the instruction mix of the sDTW cell on sm_100a.
// Measures warp-instructions per cycle per SM sub-partition (SMSP) for: FADD, FMNMX, FMNMX3,
// the cell mix (FADD, FMNMX3, FADD|abs|) and the packed variant (FADD2 for the two subtractions).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o ubench_alu ubench_alu.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CHAINS 8
#define ITERS 2048

template <int MODE> __global__ void k(float *out, long long *cycles, float seed)
{
    float a[CHAINS], b[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
        a[i] = seed + threadIdx.x * 0.001f + i;
        b[i] = seed * 0.5f + i;
    }
    float y = seed * 0.25f;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            if (MODE == 0) { // 3 FADD
                a[i] = a[i] + y; a[i] = a[i] + b[i]; a[i] = a[i] + y;
            } else if (MODE == 1) { // 3 FMNMX
                a[i] = fminf(a[i], y); a[i] = fmaxf(a[i], b[i]); a[i] = fminf(a[i], y + 1.f);
            } else if (MODE == 2) { // 3 FMNMX3
                a[i] = fminf(fminf(a[i], y), b[i]);
                a[i] = fmaxf(fmaxf(a[i], y), b[i]);
                a[i] = fminf(fminf(a[i], b[i]), y);
            } else if (MODE == 3) { // cell: t = x - y; m = min3; d = |t| + m   (chain through a[i])
                const float t = b[i] - y;
                const float m = fminf(fminf(a[i], a[(i + 1) % CHAINS]), a[(i + 2) % CHAINS]);
                a[i] = fabsf(t) + m;
            } else if (MODE == 4) { // cell with 2-input mins
                const float t = b[i] - y;
                const float m = fminf(a[i], fminf(a[(i + 1) % CHAINS], a[(i + 2) % CHAINS]));
                a[i] = fabsf(t) + m + 0.f * y;
            }
        }
        if (MODE == 5) { // cell, subtraction packed two rows at a time (add.f32x2)
#pragma unroll
            for (int i = 0; i < CHAINS; i += 2) {
                unsigned long long xb, yy, tt;
                asm("mov.b64 %0, {%1, %2};" : "=l"(xb) : "f"(b[i]), "f"(b[i + 1]));
                asm("mov.b64 %0, {%1, %1};" : "=l"(yy) : "f"(-y));
                asm("add.rn.f32x2 %0, %1, %2;" : "=l"(tt) : "l"(xb), "l"(yy));
                float t0f, t1f;
                asm("mov.b64 {%0, %1}, %2;" : "=f"(t0f), "=f"(t1f) : "l"(tt));
                const float m0 = fminf(fminf(a[i], a[(i + 1) % CHAINS]), a[(i + 2) % CHAINS]);
                a[i] = fabsf(t0f) + m0;
                const float m1 = fminf(fminf(a[i + 1], a[(i + 2) % CHAINS]), a[(i + 3) % CHAINS]);
                a[i + 1] = fabsf(t1f) + m1;
            }
        }
        y += 1e-7f;
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char *name, double instr_per_iter, int warps_per_smsp)
{
    int dev = 0;
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, dev);
    const int threads = warps_per_smsp * 4 * 32;
    const int blocks = p.multiProcessorCount;
    float *out;
    long long *cyc;
    cudaMalloc(&out, sizeof(float) * blocks * threads);
    cudaMalloc(&cyc, sizeof(long long) * blocks);
    k<MODE><<<blocks, threads>>>(out, cyc, 1.0f);
    cudaDeviceSynchronize();
    k<MODE><<<blocks, threads>>>(out, cyc, 1.0f);
    cudaDeviceSynchronize();
    long long h[1024];
    cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; i++) avg += (double)h[i];
    avg /= blocks;
    const double winstr = (double)ITERS * instr_per_iter * warps_per_smsp; // per SMSP
    printf("%-28s warps/SMSP=%2d  cycles=%.0f  warp-instr/cycle/SMSP=%.3f\n", name, warps_per_smsp, avg, winstr / avg);
    cudaFree(out);
    cudaFree(cyc);
}

int main()
{
    for (int w : {1, 2, 4, 8}) {
        run<0>("FADD x3", 3.0 * CHAINS, w);
        run<1>("FMNMX x3", 3.0 * CHAINS, w);
        run<2>("FMNMX3 x3", 3.0 * CHAINS, w);
        run<3>("cell FADD,FMNMX3,FADD", 3.0 * CHAINS, w);
        run<4>("cell 2xFMNMX (4-5 instr)", 3.0 * CHAINS, w);
        run<5>("cell FADD2-packed sub", 2.5 * CHAINS, w);
    }
    return 0;
}
