// dps_optim.cu -- the parameter-identification step on the device: what the reference's
// optimisation loop does between two simulations, as three small kernels, so that a whole identification step
// (prologue -> dps_forward -> loss -> dps_backward -> optimiser) stays on the GPU and can be captured in one CUDA graph.
//
//   dps_ident_loss     optimization_lib/ps_optimization.py:178-199 (min/max normalisation), :87-101 (default loss: mean over
//                      time of the channel-summed absolute error), :202 (nan_to_num + arg-min) and the gradient torch autograd
//                      would produce for ``loss.sum().backward()`` (:212) with respect to the recorder output
//   dps_bfgs_step      optimization_lib/optimizers.py:52-111  CustomBFGSREALOptimizer.step(): BFGS without line search, the
//                      first step scaled to the maximum relative step, every step clipped to it; one thread per trajectory
//   dps_ident_control  ps_optimization.py:246-249: the maximum step decays when the best loss of the batch did not improve
//
// All three are HBM/latency-bound bookkeeping (a few MB to a few hundred MB per call); the work is in dps_forward /
// dps_backward.  Layouts follow include/dps.h: batch index innermost.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "dps_internal.h"

namespace dps {

constexpr int OPT_MAX_PARAMS = 8;
constexpr int LOSS_THREADS = 256;

// ---------------------------------------------------------------------------------------------------------
// loss
// ---------------------------------------------------------------------------------------------------------
struct LossParams {
    long long P, Pt;
    int T, C, n_chunks, normalize;
    const double *rec, *target, *vmin, *vrange;
    double *partial, *g_rec;
};

// grid (ceil(P/256), n_chunks): thread = one trajectory, block row = one slice of the time axis.  Reads rec[c][t][p] with
// the batch index innermost (coalesced), writes the gradient in the same layout and one partial sum per (slice, p).
__global__ void __launch_bounds__(LOSS_THREADS) dps_ident_loss_kernel(const LossParams q) {
    const long long p = (long long)blockIdx.x * LOSS_THREADS + threadIdx.x;
    if (p >= q.P) return;
    const int per = (q.T + q.n_chunks - 1) / q.n_chunks;
    const int t0 = blockIdx.y * per, t1 = min(t0 + per, q.T);
    const long long pt = (q.Pt == 1) ? 0 : p;
    const double inv_T = 1.0 / (double)q.T;
    double acc = 0.0;
    for (int c = 0; c < q.C; ++c) {
        const double lo = q.normalize ? q.vmin[(long long)c * q.Pt + pt] : 0.0;
        const double rg = q.normalize ? q.vrange[(long long)c * q.Pt + pt] : 1.0;
        const double* r = q.rec + ((long long)c * q.T) * q.P + p;
        const double* y = q.target + ((long long)c * q.T) * q.Pt + pt;
        double* g = q.g_rec ? q.g_rec + ((long long)c * q.T) * q.P + p : nullptr;
#pragma unroll 4
        for (int t = t0; t < t1; ++t) {
            const double rv = r[(long long)t * q.P], yv = y[(long long)t * q.Pt];
            // (target - min)/range - (result - min)/range, as ps_optimization.py:194-195 then :99
            const double d = q.normalize ? ((yv - lo) / rg - (rv - lo) / rg) : (yv - rv);
            acc += fabs(d);
            // d|d|/d result = -sign(d) / range; mean over time: 1/T.  torch: sign(0) = 0 and sign(NaN) = 0 as well (abs'
            // backward multiplies by sgn(), which is (0 < x) - (x < 0)), so a NaN residual contributes no gradient here
            if (g) g[(long long)t * q.P] = ((d > 0.0) ? -1.0 : ((d < 0.0) ? 1.0 : 0.0)) * inv_T / rg;
        }
    }
    q.partial[(long long)blockIdx.y * q.P + p] = acc;
}

// loss[p] = (sum over slices) / T, in slice order (deterministic); block-wide arg-min of nan_to_num(loss, 1e5) per block,
// then a last-block pass over the per-block minima.  state[2] = best loss, state[3] = its row.
__global__ void __launch_bounds__(LOSS_THREADS) dps_ident_loss_finish_kernel(const LossParams q, double* loss, double* state,
                                                                           double* blk_val, long long* blk_idx,
                                                                           unsigned* done) {
    __shared__ double s_val[LOSS_THREADS];
    __shared__ long long s_idx[LOSS_THREADS];
    __shared__ bool s_last;
    const long long p = (long long)blockIdx.x * LOSS_THREADS + threadIdx.x;
    double v = INFINITY;
    long long idx = 0x7fffffffffffffffLL;
    if (p < q.P) {
        double acc = 0.0;
        for (int k = 0; k < q.n_chunks; ++k) acc += q.partial[(long long)k * q.P + p];
        acc /= (double)q.T;
        loss[p] = acc;
        v = (acc != acc) ? 100000.0 : acc;            // torch.nan_to_num(loss, 100000), ps_optimization.py:202
        idx = p;
    }
    s_val[threadIdx.x] = v;
    s_idx[threadIdx.x] = idx;
    __syncthreads();
    for (int w = LOSS_THREADS / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w) {
            const double ov = s_val[threadIdx.x + w];
            const long long oi = s_idx[threadIdx.x + w];
            if (ov < s_val[threadIdx.x] || (ov == s_val[threadIdx.x] && oi < s_idx[threadIdx.x])) {
                s_val[threadIdx.x] = ov;
                s_idx[threadIdx.x] = oi;
            }
        }
        __syncthreads();
    }
    if (state == nullptr) return;
    if (threadIdx.x == 0) {
        blk_val[blockIdx.x] = s_val[0];
        blk_idx[blockIdx.x] = s_idx[0];
        __threadfence();
        s_last = (atomicAdd(done, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double bv = INFINITY;
    long long bi = 0x7fffffffffffffffLL;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += LOSS_THREADS) {
        const double ov = __ldcg(blk_val + b);
        const long long oi = __ldcg(blk_idx + b);
        if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    s_val[threadIdx.x] = bv;
    s_idx[threadIdx.x] = bi;
    __syncthreads();
    for (int w = LOSS_THREADS / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w) {
            const double ov = s_val[threadIdx.x + w];
            const long long oi = s_idx[threadIdx.x + w];
            if (ov < s_val[threadIdx.x] || (ov == s_val[threadIdx.x] && oi < s_idx[threadIdx.x])) {
                s_val[threadIdx.x] = ov;
                s_idx[threadIdx.x] = oi;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        state[2] = s_val[0];
        state[3] = (double)s_idx[0];
        *done = 0u;                                    // ready for the next call
    }
}

static int loss_chunks(long long P, int T) {
    // enough (trajectory-block, slice) pairs to fill the GPU a few times over; at least 32 steps per slice
    int n = (int)((148LL * 8 * LOSS_THREADS + P - 1) / P);
    if (n > T / 32) n = T / 32;
    if (n < 1) n = 1;
    if (n > 64) n = 64;
    return n;
}

size_t ident_loss_workspace_bytes(long long P, int T) {
    const long long blocks = (P + LOSS_THREADS - 1) / LOSS_THREADS;
    return sizeof(double) * (size_t)(loss_chunks(P, T) * P) + (sizeof(double) + sizeof(long long)) * (size_t)blocks + 64;
}

cudaError_t launch_ident_loss(long long P, int T, int C, long long Pt, const double* rec, const double* target,
                              const double* vmin, const double* vrange, double* loss, double* g_rec, double* state,
                              void* workspace, cudaStream_t stream) {
    LossParams q;
    q.P = P; q.Pt = Pt; q.T = T; q.C = C; q.n_chunks = loss_chunks(P, T); q.normalize = (vmin != nullptr && vrange != nullptr);
    q.rec = rec; q.target = target; q.vmin = vmin; q.vrange = vrange; q.g_rec = g_rec;
    const long long blocks = (P + LOSS_THREADS - 1) / LOSS_THREADS;
    // workspace: [done counter (64 B)] [partial n_chunks x P] [blk_val blocks] [blk_idx blocks]
    unsigned* done = (unsigned*)workspace;
    q.partial = (double*)((char*)workspace + 64);
    double* blk_val = q.partial + (size_t)q.n_chunks * P;
    long long* blk_idx = (long long*)(blk_val + blocks);
    cudaError_t e = cudaMemsetAsync(done, 0, 64, stream);
    if (e != cudaSuccess) return e;
    dps_ident_loss_kernel<<<dim3((unsigned)blocks, (unsigned)q.n_chunks), LOSS_THREADS, 0, stream>>>(q);
    dps_ident_loss_finish_kernel<<<(unsigned)blocks, LOSS_THREADS, 0, stream>>>(q, loss, state, blk_val, blk_idx, done);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// BFGS step
// ---------------------------------------------------------------------------------------------------------
struct BfgsParams {
    long long P;
    int n, first_step, have_last;
    const double* state;                      // state[0] = current maximum relative step
    double* x[OPT_MAX_PARAMS];                // real part of parameter i: x[i][p * xs[i]]
    const double* g[OPT_MAX_PARAMS];          // real part of its gradient: g[i][p * gs[i]]
    int xs[OPT_MAX_PARAMS], gs[OPT_MAX_PARAMS];
    double *hesse, *last_grad, *last_sk;      // [n][n][P], [n][P], [n][P]
};

// torch.min propagates NaN (optimizers.py:76)
__device__ __forceinline__ double min_nan(double a, double b) { return (a != a || b != b) ? NAN : (a < b ? a : b); }

template <int N>
__global__ void __launch_bounds__(128) dps_bfgs_step_kernel(const BfgsParams q) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= q.P) return;
    const double max_step = q.state[0];
    double x[N], g[N], H[N][N], d[N], sk[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        x[i] = q.x[i][p * q.xs[i]];
        g[i] = q.g[i][p * q.gs[i]];
    }
    if (q.first_step) {
        // optimizers.py:64-70: hesse = eye * directions / |grads| with directions = p * max_step, so that the first step
        // equals the maximum step.  (The 0 * direction / |grad| of the off-diagonal entries is kept: it is NaN when the
        // gradient is exactly zero, as in the reference.)
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) H[i][j] = ((i == j ? 1.0 : 0.0) * (x[i] * max_step)) / fabs(g[i]);
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) H[i][j] = q.hesse[((long long)(i * N + j)) * q.P + p];
    }
    // direction = -H g (:73); a = min(1, min_i |max_step |p_i| / |d_i||) (:77-79); s = a d (:82); p += s (:83-85)
    double a = 1.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) acc += (-H[i][j]) * g[j];
        d[i] = acc;
    }
    double m = INFINITY;
#pragma unroll
    for (int i = 0; i < N; ++i) m = min_nan(m, fabs(max_step * fabs(x[i]) / fabs(d[i])));
    a = min_nan(1.0, m);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        sk[i] = a * d[i];
        q.x[i][p * q.xs[i]] = x[i] + sk[i];
    }
    if (q.have_last) {
        // :94-108: y = g - g_last; rho = 1/(y . s_last); H <- (I - rho s y^T) H (I - rho y s^T) + rho s s^T  (s = s_last)
        double y[N], s[N], dot = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            y[i] = g[i] - q.last_grad[(long long)i * q.P + p];
            s[i] = q.last_sk[(long long)i * q.P + p];
            dot += y[i] * s[i];
        }
        const double rho = 1.0 / dot;
        double HR[N][N];
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) {
                double acc = 0.0;
#pragma unroll
                for (int k = 0; k < N; ++k) acc += H[i][k] * ((k == j ? 1.0 : 0.0) - rho * (y[k] * s[j]));
                HR[i][j] = acc;
            }
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) {
                double acc = 0.0;
#pragma unroll
                for (int k = 0; k < N; ++k) acc += ((i == k ? 1.0 : 0.0) - rho * (s[i] * y[k])) * HR[k][j];
                H[i][j] = acc + rho * (s[i] * s[j]);
            }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int j = 0; j < N; ++j) q.hesse[((long long)(i * N + j)) * q.P + p] = H[i][j];
        q.last_grad[(long long)i * q.P + p] = g[i];
        q.last_sk[(long long)i * q.P + p] = sk[i];
    }
}

cudaError_t launch_bfgs_step(const BfgsParams& q, cudaStream_t stream) {
    const unsigned blocks = (unsigned)((q.P + 127) / 128);
    switch (q.n) {
#define DPS_BFGS_CASE(N) case N: dps_bfgs_step_kernel<N><<<blocks, 128, 0, stream>>>(q); break;
        DPS_BFGS_CASE(1) DPS_BFGS_CASE(2) DPS_BFGS_CASE(3) DPS_BFGS_CASE(4) DPS_BFGS_CASE(5) DPS_BFGS_CASE(6) DPS_BFGS_CASE(7)
        DPS_BFGS_CASE(8)
#undef DPS_BFGS_CASE
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// state = {max_step, last_min_loss, cur_min_loss, cur_min_row}.  ps_optimization.py:246-249:
//   if self.last_min_loss and self.last_min_loss < min_loss_val: max_step *= decay;  last_min_loss = min_loss_val
__global__ void dps_ident_control_kernel(double* state, double decay) {
    const double last = state[1], cur = state[2];
    if (last != 0.0 && last == last && last < cur) state[0] *= decay;
    state[1] = cur;
}

cudaError_t launch_ident_control(double* state, double decay, cudaStream_t stream) {
    dps_ident_control_kernel<<<1, 1, 0, stream>>>(state, decay);
    return cudaGetLastError();
}

}  // namespace dps

using namespace dps;

extern "C" {

size_t dps_ident_loss_workspace_bytes(int64_t P, int32_t T) {
    if (P < 1 || T < 1) return 0;
    return ident_loss_workspace_bytes(P, T);
}

int dps_ident_loss(int64_t P, int32_t T, int32_t C, int64_t Pt, const double* rec, const double* target,
                   const double* vmin, const double* vrange, double* loss, double* g_rec, double* state, void* workspace,
                   size_t workspace_bytes, void* stream) {
    if (P < 1 || T < 1 || C < 1 || (Pt != 1 && Pt != P)) { set_error("dps_ident_loss: bad sizes (Pt must be 1 or P)"); return 1; }
    if (!rec || !target || !loss || !workspace) { set_error("dps_ident_loss: null argument"); return 1; }
    if ((vmin == nullptr) != (vrange == nullptr)) { set_error("dps_ident_loss: vmin and vrange go together"); return 1; }
    if (workspace_bytes < ident_loss_workspace_bytes(P, T)) { set_error("dps_ident_loss: workspace too small"); return 1; }
    cudaError_t e = launch_ident_loss(P, T, C, Pt, rec, target, vmin, vrange, loss, g_rec, state, workspace, (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("dps_ident_loss: %s", cudaGetErrorString(e)); return 3; }
    return 0;
}

int dps_bfgs_step(int64_t P, int32_t n, double* const* x, const int32_t* x_stride, const double* const* grad,
                  const int32_t* grad_stride, double* hesse, double* last_grad, double* last_sk, const double* state,
                  int32_t first_step, int32_t have_last, void* stream) {
    if (P < 1 || n < 1 || n > OPT_MAX_PARAMS) { set_error("dps_bfgs_step: 1..%d parameters per trajectory", OPT_MAX_PARAMS); return 1; }
    if (!x || !x_stride || !grad || !grad_stride || !hesse || !last_grad || !last_sk || !state) { set_error("dps_bfgs_step: null argument"); return 1; }
    BfgsParams q;
    q.P = P; q.n = n; q.first_step = first_step; q.have_last = have_last; q.state = state;
    q.hesse = hesse; q.last_grad = last_grad; q.last_sk = last_sk;
    for (int i = 0; i < OPT_MAX_PARAMS; ++i) { q.x[i] = nullptr; q.g[i] = nullptr; q.xs[i] = q.gs[i] = 0; }
    for (int i = 0; i < n; ++i) {
        if (!x[i] || !grad[i] || x_stride[i] < 1 || grad_stride[i] < 1) { set_error("dps_bfgs_step: bad parameter %d", i); return 1; }
        q.x[i] = x[i]; q.g[i] = grad[i]; q.xs[i] = x_stride[i]; q.gs[i] = grad_stride[i];
    }
    cudaError_t e = launch_bfgs_step(q, (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("dps_bfgs_step: %s", cudaGetErrorString(e)); return 3; }
    return 0;
}

int dps_ident_control(double* state, double decay, void* stream) {
    if (!state) { set_error("dps_ident_control: null state"); return 1; }
    cudaError_t e = launch_ident_control(state, decay, (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("dps_ident_control: %s", cudaGetErrorString(e)); return 3; }
    return 0;
}

}  // extern "C"
