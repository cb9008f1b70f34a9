// Micro-benchmark: dependent-issue latencies seen by ONE warp of a 256-thread CTA while the other warps wait in a barrier
// (the situation of the coefficient recursion inside k_scalar).  nvcc -arch=sm_100a -O3 -fmad=false latency.cu -o latency
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long *out, double *sink, int n, int mode)
{
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    if (threadIdx.x >= 32) { if (mode & 16) __syncthreads(); return; }
    const int lane = threadIdx.x;
    double u = 1.0 + lane * 1e-6, a = 1.0000001;
    long long t0 = clock64();
    if ((mode & 15) == 0) for (int i = 0; i < n; ++i) u = fma(u, a, 1e-9);                         // dependent DFMA
    if ((mode & 15) == 1) for (int i = 0; i < n; ++i) u = __shfl_sync(0xffffffffu, u, (i + 1) & 31); // dependent SHFL (2 x 32-bit)
    if ((mode & 15) == 2) for (int i = 0; i < n; ++i) { double s = __shfl_sync(0xffffffffu, u, i & 31); u = fma(-s * a, sm[(lane * 21 + i) & 1023], u); } // one recursion step
    if ((mode & 15) == 3) for (int i = 0; i < n; ++i) u = u * a;                                  // dependent DMUL
    if ((mode & 15) == 4) for (int i = 0; i < n; ++i) u = sm[((int)u + i) & 1023];               // dependent LDS (+ cvt)
    if ((mode & 15) == 5) { // the shared-memory form of a recursion step: LDS x2, DMUL, (lane 0: STS), LDS x2, DFMA, STS, __syncwarp
        volatile double *us = sm, *rh = sm + 64, *al = sm + 128, *gs = sm + 256;
        for (int i = 0; i < n; ++i) {
            const int p = i % 10;
            const double c = rh[p] * us[p];
            if (lane == 0) al[p] = c;
            if (lane < 21) us[lane] = fma(-c * 1e-9, gs[lane * 21 + 10 + p], us[lane]);
            __syncwarp();
        }
        u += us[lane];
    }
    if ((mode & 15) == 6) for (int i = 0; i < n; ++i) { u = u / a; }                              // dependent FP64 division
    if ((mode & 15) == 7) { // STS -> LDS round trip through shared memory, same lane
        volatile double *us = sm;
        for (int i = 0; i < n; ++i) { us[lane] = u; __syncwarp(); u = us[(lane + 1) & 31] + 1e-9; __syncwarp(); }
    }
    long long t1 = clock64();
    if (lane == 0) out[0] = t1 - t0;
    sink[lane] = u;
    if (mode & 16) __syncthreads();
}
int main()
{
    long long *out; double *sink;
    cudaMalloc(&out, 8); cudaMalloc(&sink, 256);
    const char *names[] = {"dependent DFMA", "dependent SHFL(double)", "SHFL + DMUL + LDS + DFMA (one recursion step)", "dependent DMUL", "dependent LDS",
                           "shared-memory recursion step (as in compact_recursion)", "dependent FP64 division", "STS -> syncwarp -> LDS(other lane) -> DADD -> syncwarp"};
    for (int wait = 0; wait < 2; ++wait)
        for (int mode = 0; mode < 8; ++mode) {
            for (int rep = 0; rep < 2; ++rep) { k<<<1, 256>>>(out, sink, 1000, mode | (wait ? 16 : 0)); cudaDeviceSynchronize(); }
            long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
            printf("%-58s %8.1f cycles per iteration%s\n", names[mode], c / 1000.0, wait ? "  (7 other warps waiting in __syncthreads)" : "");
        }
    return 0;
}
