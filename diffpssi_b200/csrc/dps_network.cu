// dps_network.cu -- on-device fault re-reduction of the bus impedance matrix, and the fp64 peak probe.
//
// Reference behaviour being replaced (src/diffpssi/power_sim_lib/simulator.py:524-531): on EVERY step in which a
// short circuit is active the reference does two dense batched inversions,
//     Y = inv(Z);  Y[:, b, b] = 1e6;  Z = inv(Y)
// and restores Z = Z_original when it is not.  Setting one diagonal entry of Y is a rank-1 change
// dY = c e_b e_b^T with c = 1e6 - Y_bb, so by Sherman-Morrison
//     Z' = Z - f Z[:,b] Z[b,:],    f = c / (1 + c Z_bb)
// and the stepper only ever needs rows {observed buses} x columns {generator buses} of Z'.  One thread per
// trajectory (or a single thread when the network is shared by the batch) applies the updates for every
// distinct set of simultaneously faulted buses ("configuration") once, before the time loop.
#include <cuda_runtime.h>
#include "dps_internal.h"
#include "dps_reduce.cuh"
#include "dps_model.cuh"

namespace dps {

__global__ void dps_network_reduce_kernel(const ReduceParams p) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.Bz) return;
    network_reduce_one(p, b);
}

cudaError_t launch_network_reduce_params(const ReduceParams& p, cudaStream_t stream) {
    const int threads = 128;
    const long long blocks = (p.Bz + threads - 1) / threads;
    dps_network_reduce_kernel<<<(unsigned)blocks, threads, 0, stream>>>(p);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// fp64 peak probe: 8 independent DFMA chains per thread, enough warps to fill every SM
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dps_fp64_probe_kernel(double* out, int iters, double a, double b) {
    double r0 = threadIdx.x * 1e-3, r1 = r0 + 1.0, r2 = r0 + 2.0, r3 = r0 + 3.0;
    double r4 = r0 + 4.0, r5 = r0 + 5.0, r6 = r0 + 6.0, r7 = r0 + 7.0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            r0 = fma(r0, a, b); r1 = fma(r1, a, b); r2 = fma(r2, a, b); r3 = fma(r3, a, b);
            r4 = fma(r4, a, b); r5 = fma(r5, a, b); r6 = fma(r6, a, b); r7 = fma(r7, a, b);
        }
    }
    const double s = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    if (s == 12345.678) out[0] = s;      // never true: keeps the chains alive
}

// device math self-test: out[0..4][n] = sin(a), cos(a), a/b, sqrt(|a|), 1/b with the kernels' own primitives
__global__ void dps_math_probe_kernel(long long n, const double* a, const double* b, double* out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s, c;
    dps_sincos(a[i], &s, &c);
    out[i] = s;
    out[n + i] = c;
    out[2 * n + i] = dps_div(a[i], b[i]);
    out[3 * n + i] = dps_sqrt(fabs(a[i]));
    out[4 * n + i] = dps_rcp(b[i]);
}

cudaError_t launch_math_probe(long long n, const double* a, const double* b, double* out, cudaStream_t stream) {
    dps_math_probe_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(n, a, b, out);
    return cudaGetLastError();
}

cudaError_t run_fp64_probe(int iters, double* tflops, double* ms_out) {
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    double* d_out = nullptr;
    e = cudaMalloc(&d_out, sizeof(double));      // probe only (measurement helper, not the data path)
    if (e != cudaSuccess) return e;
    const int blocks = sms * 8, threads = 256;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    dps_fp64_probe_kernel<<<blocks, threads>>>(d_out, iters / 8 + 1, 0.999999, 1e-9);     // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(t0);
        dps_fp64_probe_kernel<<<blocks, threads>>>(d_out, iters, 0.999999, 1e-9);
        cudaEventRecord(t1);
        e = cudaEventSynchronize(t1);
        if (e != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t0, t1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    cudaFree(d_out);
    if (e != cudaSuccess) return e;
    const double flops = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
    *tflops = flops / (best * 1e-3) / 1e12;
    *ms_out = best;
    return cudaGetLastError();
}

}  // namespace dps
