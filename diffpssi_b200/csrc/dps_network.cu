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

// Network shared by the batch (Bz == 1): the same update, one block, threads spread over the matrix ELEMENTS.  (One
// thread walking the whole matrix with dependent global loads took 50 us -- 1 % of a bench step at 8192 trajectories
// per GPU.)  Every element of a rank-1 update only needs old values of its own row / the faulted row, so each step is
// "everybody loads, barrier, everybody stores".
__global__ void __launch_bounds__(256) dps_network_reduce_shared_kernel(const ReduceParams p) {
    const int rows = p.n_obs + p.m_max, cols = p.n_gen + p.m_max;
    const int n_el = rows * cols;
    for (int cfg = 0; cfg < p.n_cfg; ++cfg) {
        double* W = p.zsub + (long long)cfg * n_el * 2;
        const int m = p.cfg_n[cfg];
        // gather rows {observed, faulted} x columns {generators, faulted}
        for (int e = threadIdx.x; e < n_el; e += blockDim.x) {
            const int r = e / cols, c = e - r * cols;
            const int rb = (r < p.n_obs) ? p.obs_bus[r] : ((r - p.n_obs) < m ? p.cfg_bus[cfg][r - p.n_obs] : -1);
            const int cb = (c < p.n_gen) ? p.gen_bus[c] : ((c - p.n_gen) < m ? p.cfg_bus[cfg][c - p.n_gen] : -1);
            double re = 0.0, im = 0.0;
            if (rb >= 0 && cb >= 0) {
                re = p.z_re[(long long)rb * p.n_bus + cb];
                im = p.z_im[(long long)rb * p.n_bus + cb];
            }
            W[2 * e] = re;
            W[2 * e + 1] = im;
        }
        __syncthreads();
        for (int k = 0; k < m; ++k) {
            const int fb = p.cfg_bus[cfg][k];
            const int rk = p.n_obs + k, ck = p.n_gen + k;
            // f = c / (1 + c Z_bb), c = y_fault - Y_bb   (every thread computes it from the same old values)
            const double c_re = p.y_fault - p.yd_re[fb], c_im = -p.yd_im[fb];
            const double zbb_re = W[2 * (rk * cols + ck)], zbb_im = W[2 * (rk * cols + ck) + 1];
            const double den_re = 1.0 + (c_re * zbb_re - c_im * zbb_im);
            const double den_im = c_re * zbb_im + c_im * zbb_re;
            const double dn = den_re * den_re + den_im * den_im;
            const double f_re = (c_re * den_re + c_im * den_im) / dn;
            const double f_im = (c_im * den_re - c_re * den_im) / dn;
            // this thread's elements (at most a few: strided over the block); old values first ...
            constexpr int MAX_OWN = 8;
            double nre[MAX_OWN], nim[MAX_OWN];
            int own = 0;
            for (int e = threadIdx.x; e < n_el && own < MAX_OWN; e += blockDim.x, ++own) {
                const int r = e / cols, c = e - r * cols;
                const double a_re = W[2 * (r * cols + ck)], a_im = W[2 * (r * cols + ck) + 1];       // Z[r][b]
                const double b_re = W[2 * (rk * cols + c)], b_im = W[2 * (rk * cols + c) + 1];       // Z[b][c]
                const double t_re = f_re * a_re - f_im * a_im, t_im = f_re * a_im + f_im * a_re;
                nre[own] = W[2 * e] - (t_re * b_re - t_im * b_im);
                nim[own] = W[2 * e + 1] - (t_re * b_im + t_im * b_re);
            }
            __syncthreads();
            // ... then the new ones
            own = 0;
            for (int e = threadIdx.x; e < n_el && own < MAX_OWN; e += blockDim.x, ++own) {
                W[2 * e] = nre[own];
                W[2 * e + 1] = nim[own];
            }
            __syncthreads();
        }
    }
}

cudaError_t launch_network_reduce_params(const ReduceParams& p, cudaStream_t stream) {
    const int rows = p.n_obs + p.m_max, cols = p.n_gen + p.m_max;
    if (p.Bz == 1 && rows * cols <= 256 * 8) {
        dps_network_reduce_shared_kernel<<<1, 256, 0, stream>>>(p);
        return cudaGetLastError();
    }
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

// FP64 MMA probe (mma.sync.m8n8k4.f64, SASS DMMA): 8 independent accumulator pairs per warp.  On B200 the MMA shares the
// FP64 pipe with DFMA (profiles/probe_dmma_vs_dfma.txt) but issues 256 FMAs per warp instruction, so it gets closer to the
// pipe's nominal rate than a DFMA stream does; bench.py reports max(DFMA, DMMA) as the fp64 roofline denominator.
__global__ void __launch_bounds__(256) dps_fp64_dmma_probe_kernel(double* out, int iters, double a, double b) {
    double m0[8], m1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { m0[i] = threadIdx.x * 1e-3; m1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(m0[i]), "+d"(m1[i]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += m0[i] + m1[i];
    if (s == 12345.678) out[0] = s;      // never true: keeps the accumulators alive
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

cudaError_t run_fp64_dmma_probe(int iters, double* tflops, double* ms_out) {
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    double* d_out = nullptr;
    e = cudaMalloc(&d_out, sizeof(double));      // probe only (measurement helper, not the data path)
    if (e != cudaSuccess) return e;
    const int blocks = sms * 4, threads = 256;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    dps_fp64_dmma_probe_kernel<<<blocks, threads>>>(d_out, iters / 8 + 1, 0.999, 1e-3);     // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(t0);
        dps_fp64_dmma_probe_kernel<<<blocks, threads>>>(d_out, iters, 0.999, 1e-3);
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
    // one m8n8k4 MMA = 8*8*4 FMAs = 512 flops per warp instruction; 8 of them per iteration
    const double flops = 512.0 * 8.0 * (double)iters * (double)blocks * (double)(threads / 32);
    *tflops = flops / (best * 1e-3) / 1e12;
    *ms_out = best;
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
