// tests/custom_objective.cu -- USER-SIDE device objectives for the callback tests.
//
// This is what a user of lbfgsb200_create_callback() writes: plain CUDA that evaluates
// f, grad f, grad.d and grad.grad at x + alpha*d with alpha read from a device scalar.
// Built by tests/test_gpu_callback.py with nvcc into a small shared library; the product
// library knows nothing about it.
//   cb_rosenbrock : the chained Rosenbrock of parallel-implementation/functions.cpp:26-49
//                   (same expression order as the built-in, so results can be compared)
//   cb_rosenbrock_sharded : the same on one shard of a multi-GPU solver -- partial sums of the shard, neighbours'
//                   boundary values from lbfgsb200_device_halo()
//   cb_rosenbrock_syncing : cb_rosenbrock + a read-back of f to the host (copy + cudaStreamSynchronize): legal in the
//                   host-stepped loop, impossible to capture -- the solver must fall back by itself
//   cb_dense      : f = x^T A x + b^T x with a dense SPD A (the reference's unused fixtures,
//                   sequential-implementation/matrices.h)
#include <cuda_runtime.h>
#include <stddef.h>

struct DenseCtx {
    const double *A; // n x n row-major, device
    const double *b; // n, device
    double *tmp;     // 3n doubles of device scratch
};

// out3 = { sum t0, sum t1, sum t2 } over n elements, one block, fixed order
__global__ void reduce3(const double *t0, const double *t1, const double *t2, size_t n, double *out3)
{
    __shared__ double sm[3][256];
    double a = 0, b = 0, c = 0;
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) {
        a += t0[i];
        b += t1[i];
        c += t2[i];
    }
    sm[0][threadIdx.x] = a; sm[1][threadIdx.x] = b; sm[2][threadIdx.x] = c;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s)
            for (int q = 0; q < 3; ++q) sm[q][threadIdx.x] += sm[q][threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) { out3[0] = sm[0][0]; out3[1] = sm[1][0]; out3[2] = sm[2][0]; }
}

__global__ void rosen_eval(const double *x, const double *d, const double *d_alpha, double *g, double *tf, double *tgd,
                           double *tgg, size_t n)
{
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double a = *d_alpha;
    const double c = x[i] + a * d[i];
    const double l = i > 0 ? x[i - 1] + a * d[i - 1] : 0.0;
    const double r = i + 1 < n ? x[i + 1] + a * d[i + 1] : 0.0;
    const double bl = c - l * l, bc = r - c * c, t2 = 1 - c;
    const bool hl = i > 0, hr = i + 1 < n;
    const double from_left = hl ? 200.0 * bl : 0.0;
    const double gv = hr ? from_left + (2.0 * (c - 1) - 400.0 * c * bc) : from_left;
    g[i] = gv;
    tf[i] = hr ? 100.0 * bc * bc + t2 * t2 : 0.0;
    tgd[i] = gv * d[i];
    tgg[i] = gv * gv;
}

// one shard [goff, goff + n) of a chain of nglob elements; halo = { xL, xR, dL, dR, .. } (device, the solver's)
__global__ void rosen_eval_sharded(const double *x, const double *d, const double *d_alpha, double *g, double *tf, double *tgd,
                                   double *tgg, size_t n, size_t goff, size_t nglob, const double *halo)
{
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double a = *d_alpha;
    const bool hl = goff + i > 0, hr = goff + i + 1 < nglob;
    const double c = x[i] + a * d[i];
    const double l = i > 0 ? x[i - 1] + a * d[i - 1] : (hl ? halo[0] + a * halo[2] : 0.0);
    const double r = i + 1 < n ? x[i + 1] + a * d[i + 1] : (hr ? halo[1] + a * halo[3] : 0.0);
    const double bl = c - l * l, bc = r - c * c, t2 = 1 - c;
    const double from_left = hl ? 200.0 * bl : 0.0;
    const double gv = hr ? from_left + (2.0 * (c - 1) - 400.0 * c * bc) : from_left;
    g[i] = gv;
    tf[i] = hr ? 100.0 * bc * bc + t2 * t2 : 0.0;
    tgd[i] = gv * d[i];
    tgg[i] = gv * gv;
}

struct ShardCtx {
    double *tmp;        // 3 n_local doubles of device scratch
    const double *halo; // lbfgsb200_device_halo(solver), filled in after create
    size_t n_global;
};

__global__ void dense_eval(const double *A, const double *b, const double *x, const double *d, const double *d_alpha,
                           double *g, double *tf, double *tgd, double *tgg, size_t n)
{
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double a = *d_alpha;
    double ax = 0.0, atx = 0.0;
    for (size_t j = 0; j < n; ++j) {
        const double xj = x[j] + a * d[j];
        ax += A[i * n + j] * xj;
        atx += A[j * n + i] * xj;
    }
    const double xi = x[i] + a * d[i];
    const double gv = ax + atx + b[i];
    g[i] = gv;
    tf[i] = xi * ax + b[i] * xi;
    tgd[i] = gv * d[i];
    tgg[i] = gv * gv;
}

extern "C" {

// user = device scratch of 3n doubles
int cb_rosenbrock(const double *x, const double *d, const double *d_alpha, double *g_out, double *d_out3, size_t n,
                  size_t global_offset, void *user, void *stream)
{
    (void)global_offset;
    double *tmp = (double *)user;
    cudaStream_t st = (cudaStream_t)stream;
    rosen_eval<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, d, d_alpha, g_out, tmp, tmp + n, tmp + 2 * n, n);
    reduce3<<<1, 256, 0, st>>>(tmp, tmp + n, tmp + 2 * n, n, d_out3);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int cb_rosenbrock_sharded(const double *x, const double *d, const double *d_alpha, double *g_out, double *d_out3, size_t n,
                          size_t global_offset, void *user, void *stream)
{
    const ShardCtx *c = (const ShardCtx *)user;
    cudaStream_t st = (cudaStream_t)stream;
    rosen_eval_sharded<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, d, d_alpha, g_out, c->tmp, c->tmp + n, c->tmp + 2 * n, n,
                                                                    global_offset, c->n_global, c->halo);
    reduce3<<<1, 256, 0, st>>>(c->tmp, c->tmp + n, c->tmp + 2 * n, n, d_out3);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int cb_rosenbrock_syncing(const double *x, const double *d, const double *d_alpha, double *g_out, double *d_out3, size_t n,
                          size_t global_offset, void *user, void *stream)
{
    // a callback that looks at its result on the host (logging, say): fine in the host-stepped loop, impossible to
    // record -- the synchronisation fails while the stream is being captured and the callback reports it
    if (cb_rosenbrock(x, d, d_alpha, g_out, d_out3, n, global_offset, user, stream)) return 1;
    double f_host = 0.0;
    cudaError_t e = cudaMemcpyAsync(&f_host, d_out3, sizeof(double), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    cudaGetLastError();
    return e == cudaSuccess && f_host == f_host ? 0 : 1;
}

int cb_dense(const double *x, const double *d, const double *d_alpha, double *g_out, double *d_out3, size_t n,
             size_t global_offset, void *user, void *stream)
{
    (void)global_offset;
    const DenseCtx *c = (const DenseCtx *)user;
    cudaStream_t st = (cudaStream_t)stream;
    dense_eval<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(c->A, c->b, x, d, d_alpha, g_out, c->tmp, c->tmp + n, c->tmp + 2 * n, n);
    reduce3<<<1, 256, 0, st>>>(c->tmp, c->tmp + n, c->tmp + 2 * n, n, d_out3);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int cb_always_fails(const double *, const double *, const double *, double *, double *, size_t, size_t, void *, void *)
{
    return 1;
}

} // extern "C"
