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
    long long t1 = clock64();
    if (lane == 0) out[0] = t1 - t0;
    sink[lane] = u;
    if (mode & 16) __syncthreads();
}
int main()
{
    long long *out; double *sink;
    cudaMalloc(&out, 8); cudaMalloc(&sink, 256);
    const char *names[] = {"dependent DFMA", "dependent SHFL(double)", "SHFL + DMUL + LDS + DFMA (one recursion step)", "dependent DMUL", "dependent LDS"};
    for (int mode = 0; mode < 5; ++mode) {
        for (int rep = 0; rep < 2; ++rep) { k<<<1, 256>>>(out, sink, 1000, mode); cudaDeviceSynchronize(); }
        long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
        printf("%-50s %8.1f cycles per iteration\n", names[mode], c / 1000.0);
    }
    return 0;
}
