// dps_prologue.cu -- initialize() on the device: what PowerSystemSimulation.run() does BEFORE the
// time loop, as three kernels behind the C ABI instead of ~2000 eager tensor operations:
//
//   dps_load_flow        polar Newton-Raphson power flow, one block per batch row        load_flow.py:9-179
//   dps_dynamic_inverse  loads -> constant admittances from the load-flow voltage,       static_models.py:515-527,
//                        Y_dyn = base + loads, Z0 = inv(Y_dyn) (complex Gauss-Jordan)    simulator.py:386, 400-422
//   dps_model_init       initial states, derived constants ("slots") and Norton          synchronous_machine.py:279-323,
//                        currents of every (trajectory, generator)                        exciters.py:129-138, governors.py:136-145,
//                                                                                         stabilizers.py:135-145, blocks.py:196-207
//
// A batch row is a trajectory (B = parallel_sims) or, when the network is shared by the batch, the single row B = 1: the
// reference's O(P^2) initialisation (index tensors scattering with duplicates, simulator.py:367-380) has no counterpart.
// The host side (diffpssi_b200/prologue_device.py) uses these entry points whenever no input carries an autograd graph;
// runs that need gradients through initialize() keep the differentiable torch prologue (prologue.py).
// Complex arrays are interleaved (re, im) pairs -- torch.complex128 memory -- with the batch index OUTERMOST, as torch
// produces them; per-generator outputs are structure-of-arrays with the batch index innermost (include/dps.h).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "dps_internal.h"

namespace dps {

constexpr int PRO_MAX_BUS = 64;          // dense Jacobian / Gauss-Jordan workspaces live in shared memory
constexpr int LF_THREADS = 128;
constexpr int INV_THREADS = 256;

struct LfParams {
    int n, n_ang, n_pq;
    long long B, By, Bs;
    const double2* y;        // [By][n][n]
    const double2* s_soll;   // [Bs][n]
    double2* v;              // [B][n] current iterate, updated in place
    double2* s_out;          // [B][n]
    double* worst;           // [1] max |mismatch| over the batch (pre-update, like load_flow.py:147)
    unsigned char ang[PRO_MAX_BUS];   // bus of angle unknown r (PV buses, then PQ buses)
    unsigned char pq[PRO_MAX_BUS];    // bus of magnitude unknown r
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cdiv(double2 a, double2 b) {
    const double d = b.x * b.x + b.y * b.y;
    return make_double2((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d);
}

// S_i = V_i conj(sum_k Y_ik V_k) for every bus of one batch row (threads < n)
__device__ __forceinline__ double2 bus_power(const double2* Y, const double2* V, int n, int i) {
    double2 acc = make_double2(0.0, 0.0);
    for (int k = 0; k < n; ++k) {
        const double2 y = Y[i * n + k], v = V[k];
        acc.x += y.x * v.x - y.y * v.y;
        acc.y += y.x * v.y + y.y * v.x;
    }
    return cmul(V[i], make_double2(acc.x, -acc.y));
}

// One Newton-Raphson iteration of one batch row: mismatch, Jacobian (minus the usual one, as load_flow.py:75-140 builds
// it), dense solve with partial pivoting, polar update (angles of PV and PQ buses, then magnitudes of PQ buses).
__global__ void __launch_bounds__(LF_THREADS) dps_lf_iter_kernel(const LfParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = p.n, m = p.n_ang + p.n_pq, ld = m + 1;
    double2* V = reinterpret_cast<double2*>(smem_raw);      // [n]
    double* th = reinterpret_cast<double*>(V + n);           // [n]
    double* vm = th + n;                                     // [n]
    double* Pc = vm + n;                                     // [n]
    double* Qc = Pc + n;                                     // [n]
    double* xs = Qc + n;                                     // [m] solution
    double* red = xs + m;                                    // [LF_THREADS] reduction scratch
    int* redi = reinterpret_cast<int*>(red + LF_THREADS);    // [LF_THREADS]
    double* A = reinterpret_cast<double*>(redi + LF_THREADS);// [m][m+1] augmented system (-J | rhs)
    const int t = threadIdx.x;
    const long long b = blockIdx.x;
    const double2* Y = p.y + (p.By > 1 ? b : 0) * (long long)n * n;
    const double2* S0 = p.s_soll + (p.Bs > 1 ? b : 0) * (long long)n;
    double2* Vg = p.v + b * n;

    for (int i = t; i < n; i += LF_THREADS) V[i] = Vg[i];
    __syncthreads();
    for (int i = t; i < n; i += LF_THREADS) {
        const double2 s = bus_power(Y, V, n, i);
        Pc[i] = s.x;
        Qc[i] = s.y;
        th[i] = atan2(V[i].y, V[i].x);
        vm[i] = hypot(V[i].x, V[i].y);
    }
    __syncthreads();
    double wmax = 0.0;
    for (int idx = t; idx < m * ld; idx += LF_THREADS) {
        const int r = idx / ld, c = idx - r * ld;
        const bool r_ang = r < p.n_ang;
        const int i = r_ang ? p.ang[r] : p.pq[r - p.n_ang];
        if (c == m) {                                        // right-hand side: P mismatch of angle rows, Q mismatch of PQ rows
            const double rhs = r_ang ? (Pc[i] - S0[i].x) : (Qc[i] - S0[i].y);
            A[idx] = rhs;
            wmax = fmax(wmax, fabs(rhs));
            continue;
        }
        const bool c_ang = c < p.n_ang;
        const int k = c_ang ? p.ang[c] : p.pq[c - p.n_ang];
        const double G = Y[i * n + k].x, Bm = Y[i * n + k].y;
        double J;
        if (i != k) {
            double sn, cs;
            sincos(th[i] - th[k], &sn, &cs);
            const double a = G * sn - Bm * cs, cc = G * cs + Bm * sn;
            if (r_ang) J = c_ang ? -vm[i] * vm[k] * a : -vm[i] * cc;          // J11 / J12
            else J = c_ang ? vm[i] * vm[k] * cc : -vm[i] * a;                  // J21 / J22
        } else {
            if (r_ang) J = c_ang ? (vm[i] * vm[i] * Bm + Qc[i]) : (-vm[i] * G - Pc[i] / vm[i]);
            else J = c_ang ? (vm[i] * vm[i] * G - Pc[i]) : (vm[i] * Bm - Qc[i] / vm[i]);
        }
        A[idx] = -J;
    }
    red[t] = wmax;
    __syncthreads();
    for (int s = LF_THREADS / 2; s > 0; s >>= 1) {
        if (t < s) red[t] = fmax(red[t], red[t + s]);
        __syncthreads();
    }
    // non-negative doubles order like their bit patterns
    if (t == 0) atomicMax(reinterpret_cast<unsigned long long*>(p.worst), (unsigned long long)__double_as_longlong(red[0]));
    __syncthreads();

    // Gaussian elimination with partial pivoting on the augmented matrix
    for (int piv = 0; piv < m; ++piv) {
        double best = -1.0;
        int arg = piv;
        for (int r = piv + t; r < m; r += LF_THREADS) {
            const double v = fabs(A[r * ld + piv]);
            if (v > best) { best = v; arg = r; }
        }
        red[t] = best;
        redi[t] = arg;
        __syncthreads();
        for (int s = LF_THREADS / 2; s > 0; s >>= 1) {
            if (t < s && (red[t + s] > red[t] || (red[t + s] == red[t] && redi[t + s] < redi[t]))) {
                red[t] = red[t + s];
                redi[t] = redi[t + s];
            }
            __syncthreads();
        }
        const int pr = redi[0];
        __syncthreads();
        if (pr != piv)
            for (int c = piv + t; c < ld; c += LF_THREADS) {
                const double tmp = A[piv * ld + c];
                A[piv * ld + c] = A[pr * ld + c];
                A[pr * ld + c] = tmp;
            }
        __syncthreads();
        const double inv = 1.0 / A[piv * ld + piv];
        for (int r = piv + 1 + t; r < m; r += LF_THREADS) xs[r] = A[r * ld + piv] * inv;     // multipliers
        __syncthreads();
        const int w = ld - piv - 1;                                                           // columns piv+1 .. m
        for (int idx = t; idx < (m - piv - 1) * w; idx += LF_THREADS) {
            const int r = piv + 1 + idx / w, c = piv + 1 + idx % w;
            A[r * ld + c] = fma(-xs[r], A[piv * ld + c], A[r * ld + c]);
        }
        __syncthreads();
    }
    for (int c = m - 1; c >= 0; --c) {
        if (t == 0) xs[c] = A[c * ld + m] / A[c * ld + c];
        __syncthreads();
        for (int r = t; r < c; r += LF_THREADS) A[r * ld + m] = fma(-A[r * ld + c], xs[c], A[r * ld + m]);
        __syncthreads();
    }
    // update (load_flow.py:150-158)
    for (int r = t; r < m; r += LF_THREADS) {
        if (r < p.n_ang) th[p.ang[r]] -= xs[r];
        else vm[p.pq[r - p.n_ang]] -= xs[r];
    }
    __syncthreads();
    for (int i = t; i < n; i += LF_THREADS) {
        double sn, cs;
        sincos(th[i], &sn, &cs);
        Vg[i] = make_double2(vm[i] * cs, vm[i] * sn);
    }
}

__global__ void dps_lf_power_kernel(const LfParams p) {
    const long long b = blockIdx.x;
    const int n = p.n;
    const double2* Y = p.y + (p.By > 1 ? b : 0) * (long long)n * n;
    const double2* V = p.v + b * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) p.s_out[b * n + i] = bus_power(Y, V, n, i);
}

__global__ void dps_lf_seed_kernel(double2* v, const double2* v_init, long long Bv, int n, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) v[i] = v_init[(Bv > 1) ? i : (i % n)];
}

struct InvParams {
    int n, n_load;
    long long B, By, Bl, Bv;
    const double2* y_base;   // [By][n][n]
    const double2* s_load;   // [Bl][n_load]
    const double2* v;        // [Bv][n]
    double2* y_load;         // [B][n_load]
    double2* y_diag;         // [B][n]
    double2* z0;             // [B][n][n]
    unsigned char load_bus[PRO_MAX_BUS * 2];
};

// Y_dyn = base + diag(loads as constant admittances), Z0 = inv(Y_dyn): Gauss-Jordan with partial pivoting on [Y | I]
__global__ void __launch_bounds__(INV_THREADS) dps_dyn_inverse_kernel(const InvParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = p.n, ld = 2 * n;
    double2* M = reinterpret_cast<double2*>(smem_raw);        // [n][2n]
    double2* col = M + n * ld;                                // [n] pivot column
    double* red = reinterpret_cast<double*>(col + n);         // [INV_THREADS]
    int* redi = reinterpret_cast<int*>(red + INV_THREADS);
    const int t = threadIdx.x;
    const long long b = blockIdx.x;
    const double2* Y = p.y_base + (p.By > 1 ? b : 0) * (long long)n * n;
    const double2* V = p.v + (p.Bv > 1 ? b : 0) * (long long)n;
    for (int idx = t; idx < n * ld; idx += INV_THREADS) {
        const int r = idx / ld, c = idx - r * ld;
        M[idx] = (c < n) ? Y[r * n + c] : make_double2((c - n == r) ? 1.0 : 0.0, 0.0);
    }
    __syncthreads();
    if (t == 0) {
        // static_models.py:525-527: z = conj(|v|^2 / s_load), y = 1 / z
        for (int l = 0; l < p.n_load; ++l) {
            const int bus = p.load_bus[l];
            const double va = hypot(V[bus].x, V[bus].y);
            const double2 q = cdiv(make_double2(va * va, 0.0), p.s_load[(p.Bl > 1 ? b : 0) * (long long)p.n_load + l]);
            const double2 y = cdiv(make_double2(1.0, 0.0), make_double2(q.x, -q.y));
            p.y_load[b * p.n_load + l] = y;
            M[bus * ld + bus].x += y.x;
            M[bus * ld + bus].y += y.y;
        }
    }
    __syncthreads();
    for (int i = t; i < n; i += INV_THREADS) p.y_diag[b * n + i] = M[i * ld + i];
    for (int piv = 0; piv < n; ++piv) {
        double best = -1.0;
        int arg = piv;
        for (int r = piv + t; r < n; r += INV_THREADS) {
            const double2 e = M[r * ld + piv];
            const double v = e.x * e.x + e.y * e.y;
            if (v > best) { best = v; arg = r; }
        }
        red[t] = best;
        redi[t] = arg;
        __syncthreads();
        for (int s = INV_THREADS / 2; s > 0; s >>= 1) {
            if (t < s && (red[t + s] > red[t] || (red[t + s] == red[t] && redi[t + s] < redi[t]))) {
                red[t] = red[t + s];
                redi[t] = redi[t + s];
            }
            __syncthreads();
        }
        const int pr = redi[0];
        __syncthreads();
        if (pr != piv)
            for (int c = t; c < ld; c += INV_THREADS) {
                const double2 tmp = M[piv * ld + c];
                M[piv * ld + c] = M[pr * ld + c];
                M[pr * ld + c] = tmp;
            }
        __syncthreads();
        const double2 inv = cdiv(make_double2(1.0, 0.0), M[piv * ld + piv]);
        for (int r = t; r < n; r += INV_THREADS) col[r] = M[r * ld + piv];
        __syncthreads();
        for (int c = t; c < ld; c += INV_THREADS) M[piv * ld + c] = cmul(M[piv * ld + c], inv);
        __syncthreads();
        for (int idx = t; idx < n * ld; idx += INV_THREADS) {
            const int r = idx / ld, c = idx - r * ld;
            if (r == piv) continue;
            const double2 f = col[r], e = M[piv * ld + c];
            M[idx].x -= f.x * e.x - f.y * e.y;
            M[idx].y -= f.x * e.y + f.y * e.x;
        }
        __syncthreads();
    }
    double2* Z = p.z0 + b * (long long)n * n;
    for (int idx = t; idx < n * n; idx += INV_THREADS) {
        const int r = idx / n, c = idx - r * n;
        Z[idx] = M[r * ld + n + c];
    }
}

struct InitParams {
    int n_gen, n_bus;
    long long B, Bv, Bs;
    const double2* v;        // [Bv][n_bus] load-flow voltage
    const double2* s;        // [Bs][n_bus] load-flow power
    const double* raw_u;     // [n_gen][DPS_N_RAW]
    const double* raw_v;     // [n_rv][B]
    double* x0;              // [n_gen][13][B]
    double* slots;           // [n_gen][42][B]
    double* inj;             // [n_gen][2][B]
    short gen_bus[MAX_GW];
    unsigned char ctrl[MAX_GW];
    short raw_vary[MAX_GW][DPS_N_RAW];
};

__device__ __forceinline__ double clamp_mm(double x, double lo, double hi) { return fmin(fmax(x, lo), hi); }

// One thread per (trajectory, generator): synchronous_machine.py:279-323 and the controller initialisations.  States are
// real (their imaginary parts are identically zero in the reference).
__global__ void dps_model_init_kernel(const InitParams p) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int g = blockIdx.y;
    if (b >= p.B) return;
    double r[DPS_N_RAW];
#pragma unroll
    for (int k = 0; k < DPS_N_RAW; ++k) {
        const int vi = p.raw_vary[g][k];
        r[k] = (vi >= 0) ? p.raw_v[(long long)vi * p.B + b] : p.raw_u[g * DPS_N_RAW + k];
    }
    const int bus = p.gen_bus[g], ctrl = p.ctrl[g];
    const double2 v = p.v[(p.Bv > 1 ? b : 0) * (long long)p.n_bus + bus];
    const double2 sb = p.s[(p.Bs > 1 ? b : 0) * (long long)p.n_bus + bus];
    const double snr = r[DPS_RAW_SN] / r[DPS_RAW_SN_SYS];
    // i = conj(S / v) / (S_n / S_sys);  e = v + j x_q i;  delta = angle(e)
    double2 i_gen = cdiv(sb, v);
    i_gen = make_double2(i_gen.x / snr, -i_gen.y / snr);
    const double2 e = make_double2(v.x - r[DPS_RAW_XQ] * i_gen.y, v.y + r[DPS_RAW_XQ] * i_gen.x);
    const double delta = atan2(e.y, e.x);
    double sn, cs;
    sincos(1.5707963267948966 - delta, &sn, &cs);                // rot = exp(j (pi/2 - delta))
    const double2 rot = make_double2(cs, sn);
    const double2 i_dq = cmul(i_gen, rot), v_dq = cmul(v, rot);
    const double i_d = i_dq.x, i_q = i_dq.y, v_d = v_dq.x, v_q = v_dq.y;
    const double p_m = sb.x / snr;
    const double e_q_t = v_q + r[DPS_RAW_XDT] * i_d;
    const double e_d_t = v_d - r[DPS_RAW_XQT] * i_q;
    const double e_q_st = v_q + r[DPS_RAW_XDST] * i_d;
    const double e_d_st = v_d - r[DPS_RAW_XQST] * i_q;
    const double e_fd = e_q_t + i_d * (r[DPS_RAW_XD] - r[DPS_RAW_XDT]);

    double x[DPS_N_STATES], sl[DPS_N_SLOTS];
#pragma unroll
    for (int k = 0; k < DPS_N_STATES; ++k) x[k] = 0.0;
#pragma unroll
    for (int k = 0; k < DPS_N_SLOTS; ++k) sl[k] = 0.0;
    x[DPS_X_DELTA] = delta; x[DPS_X_EQT] = e_q_t; x[DPS_X_EDT] = e_d_t; x[DPS_X_EQST] = e_q_st; x[DPS_X_EDST] = e_d_st;
    sl[DPS_S_INV_2H] = 1.0 / (2.0 * r[DPS_RAW_H]);
    sl[DPS_S_D] = r[DPS_RAW_D];
    sl[DPS_S_INV_TD0T] = 1.0 / r[DPS_RAW_TD0T];
    sl[DPS_S_XD_XDT] = r[DPS_RAW_XD] - r[DPS_RAW_XDT];
    sl[DPS_S_INV_TQ0T] = 1.0 / r[DPS_RAW_TQ0T];
    sl[DPS_S_XQ_XQT] = r[DPS_RAW_XQ] - r[DPS_RAW_XQT];
    sl[DPS_S_INV_TD0ST] = 1.0 / r[DPS_RAW_TD0ST];
    sl[DPS_S_XDT_XDST] = r[DPS_RAW_XDT] - r[DPS_RAW_XDST];
    sl[DPS_S_INV_TQ0ST] = 1.0 / r[DPS_RAW_TQ0ST];
    sl[DPS_S_XQT_XQST] = r[DPS_RAW_XQT] - r[DPS_RAW_XQST];
    sl[DPS_S_INV_XDST] = 1.0 / r[DPS_RAW_XDST];
    sl[DPS_S_INV_XQST] = 1.0 / r[DPS_RAW_XQST];
    sl[DPS_S_SNR] = snr;
    sl[DPS_S_PM0] = p_m;
    sl[DPS_S_EFD0] = e_fd;
    if (ctrl & DPS_CTRL_AVR) {
        // exciters.py:137-138, blocks.py:206-207, 341-342: lead-lag state = bias = e_fd / K, PT1 state = clamp(e_fd)
        const double bias = e_fd / r[DPS_RAW_A_K];
        x[DPS_X_AVR_LL] = bias;
        x[DPS_X_AVR_PT1] = clamp_mm(e_fd, r[DPS_RAW_A_EMIN], r[DPS_RAW_A_EMAX]);
        const double ratio = r[DPS_RAW_A_T1] / r[DPS_RAW_A_T2];
        sl[DPS_S_A_VSET] = r[DPS_RAW_A_VSET];
        sl[DPS_S_A_BIAS] = bias;
        sl[DPS_S_A_R] = ratio;
        sl[DPS_S_A_OMR] = 1.0 - ratio;
        sl[DPS_S_A_INV_TB] = 1.0 / r[DPS_RAW_A_T2];
        sl[DPS_S_A_INV_TE] = 1.0 / r[DPS_RAW_A_TE];
        sl[DPS_S_A_K] = r[DPS_RAW_A_K];
        sl[DPS_S_A_EMIN] = r[DPS_RAW_A_EMIN];
        sl[DPS_S_A_EMAX] = r[DPS_RAW_A_EMAX];
    }
    if (ctrl & DPS_CTRL_GOV) {
        // governors.py:144-145: PT1 state = clamp(p_m), lead-lag state = p_m, p_ref = p_m / gain
        x[DPS_X_GOV_PT1] = clamp_mm(p_m, r[DPS_RAW_G_VMIN], r[DPS_RAW_G_VMAX]);
        x[DPS_X_GOV_LL] = p_m;
        const double ratio = r[DPS_RAW_G_T2] / r[DPS_RAW_G_T3];
        sl[DPS_S_G_PREF] = p_m / r[DPS_RAW_G_GAIN];
        sl[DPS_S_G_INV_T1] = 1.0 / r[DPS_RAW_G_T1];
        sl[DPS_S_G_GAIN] = r[DPS_RAW_G_GAIN];
        sl[DPS_S_G_VMIN] = r[DPS_RAW_G_VMIN];
        sl[DPS_S_G_VMAX] = r[DPS_RAW_G_VMAX];
        sl[DPS_S_G_R] = ratio;
        sl[DPS_S_G_OMR] = 1.0 - ratio;
        sl[DPS_S_G_INV_T3] = 1.0 / r[DPS_RAW_G_T3];
        sl[DPS_S_G_DT] = r[DPS_RAW_G_DT];
    }
    if (ctrl & DPS_CTRL_PSS) {
        // synchronous_machine.py:323: the stabiliser starts from e_fd * 0 -> all three states zero
        const double r1 = r[DPS_RAW_S_LL1_T1] / r[DPS_RAW_S_LL1_T2], r2 = r[DPS_RAW_S_LL2_T1] / r[DPS_RAW_S_LL2_T2];
        sl[DPS_S_S_INV_TW] = 1.0 / r[DPS_RAW_S_TW];
        sl[DPS_S_S_KW] = r[DPS_RAW_S_KW];
        sl[DPS_S_S_R1] = r1;
        sl[DPS_S_S_OMR1] = 1.0 - r1;
        sl[DPS_S_S_INV_T3] = 1.0 / r[DPS_RAW_S_LL1_T2];
        sl[DPS_S_S_R2] = r2;
        sl[DPS_S_S_OMR2] = 1.0 - r2;
        sl[DPS_S_S_INV_T4] = 1.0 / r[DPS_RAW_S_LL2_T2];
        sl[DPS_S_S_HLIM] = r[DPS_RAW_S_HLIM];
    }
#pragma unroll
    for (int k = 0; k < DPS_N_STATES; ++k) p.x0[((long long)g * DPS_N_STATES + k) * p.B + b] = x[k];
#pragma unroll
    for (int k = 0; k < DPS_N_SLOTS; ++k) p.slots[((long long)g * DPS_N_SLOTS + k) * p.B + b] = sl[k];
    // Norton current at x0 (synchronous_machine.py:256-260): I = snr [(a s - b c) - j (a c + b s)]
    double sd, cd;
    sincos(delta, &sd, &cd);
    const double a = e_q_st / r[DPS_RAW_XDST], bq = e_d_st / r[DPS_RAW_XQST];
    p.inj[((long long)g * 2 + 0) * p.B + b] = snr * (a * sd - bq * cd);
    p.inj[((long long)g * 2 + 1) * p.B + b] = -snr * (a * cd + bq * sd);
}

}  // namespace dps

using namespace dps;

extern "C" {

int dps_load_flow(int32_t n_bus, const int32_t* bus_type, int64_t B, const double* y_static, int64_t By,
                  const double* s_soll, int64_t Bs, const double* v_init, int64_t Bv, double* v_out, double* s_out,
                  double* worst, double tol, int32_t max_iter, int32_t* iters, void* stream) {
    if (n_bus < 1 || n_bus > PRO_MAX_BUS) { set_error("dps_load_flow: %d buses (the device load flow handles 1..%d)", n_bus, PRO_MAX_BUS); return 2; }
    if (B < 1 || !bus_type || !y_static || !s_soll || !v_init || !v_out || !s_out || !worst || !iters) { set_error("dps_load_flow: bad argument"); return 1; }
    if ((By != 1 && By != B) || (Bs != 1 && Bs != B) || (Bv != 1 && Bv != B)) { set_error("dps_load_flow: batch sizes must be 1 or B"); return 1; }
    cudaStream_t st = (cudaStream_t)stream;
    LfParams p;
    memset(&p, 0, sizeof(p));
    p.n = n_bus; p.B = B; p.By = By; p.Bs = Bs;
    p.y = (const double2*)y_static; p.s_soll = (const double2*)s_soll; p.v = (double2*)v_out; p.s_out = (double2*)s_out; p.worst = worst;
    // unknown ordering of load_flow.py:37-60: angles of PV buses then PQ buses, magnitudes of PQ buses
    for (int i = 0; i < n_bus; ++i) if (bus_type[i] == DPS_BUS_PV) p.ang[p.n_ang++] = (unsigned char)i;
    for (int i = 0; i < n_bus; ++i) if (bus_type[i] == DPS_BUS_PQ) { p.ang[p.n_ang++] = (unsigned char)i; p.pq[p.n_pq++] = (unsigned char)i; }
    const int m = p.n_ang + p.n_pq;
    const long long total = B * n_bus;
    dps_lf_seed_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((double2*)v_out, (const double2*)v_init, Bv, n_bus, total);
    const size_t smem = sizeof(double2) * n_bus + sizeof(double) * (4 * n_bus + m + LF_THREADS) + sizeof(int) * LF_THREADS +
                        sizeof(double) * (size_t)m * (m + 1);
    cudaError_t e = cudaFuncSetAttribute(dps_lf_iter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("dps_load_flow: %s", cudaGetErrorString(e)); return 3; }
    int it = 0;
    if (m > 0) {
        for (;;) {
            double w = 0.0;
            cudaMemsetAsync(worst, 0, sizeof(double), st);
            dps_lf_iter_kernel<<<(unsigned)B, LF_THREADS, smem, st>>>(p);
            // the reference's stopping rule is a host decision on the batch-global mismatch (load_flow.py:147-167): one
            // device->host read per Newton iteration, as its own loop has
            e = cudaMemcpyAsync(&w, worst, sizeof(double), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { set_error("dps_load_flow: %s", cudaGetErrorString(e)); return 3; }
            if (!(w >= tol)) { if (w != w) { set_error("dps_load_flow: NaN mismatch"); return 4; } break; }
            if (it > max_iter) { set_error("Load flow did not converge"); return 4; }
            ++it;
        }
    }
    dps_lf_power_kernel<<<(unsigned)B, 64, 0, st>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("dps_load_flow: %s", cudaGetErrorString(e)); return 3; }
    *iters = it;
    return 0;
}

int dps_dynamic_inverse(int32_t n_bus, int64_t B, const double* y_base, int64_t By, int32_t n_load, const int32_t* load_bus,
                        const double* s_load, int64_t Bl, const double* v, int64_t Bv, double* y_load, double* y_diag,
                        double* z0, void* stream) {
    if (n_bus < 1 || n_bus > PRO_MAX_BUS) { set_error("dps_dynamic_inverse: %d buses (the device inverse handles 1..%d)", n_bus, PRO_MAX_BUS); return 2; }
    if (n_load < 0 || n_load > 2 * PRO_MAX_BUS) { set_error("dps_dynamic_inverse: too many loads (%d)", n_load); return 2; }
    if (B < 1 || !y_base || !v || !y_diag || !z0 || (n_load > 0 && (!load_bus || !s_load || !y_load))) { set_error("dps_dynamic_inverse: bad argument"); return 1; }
    if ((By != 1 && By != B) || (Bl != 1 && Bl != B) || (Bv != 1 && Bv != B)) { set_error("dps_dynamic_inverse: batch sizes must be 1 or B"); return 1; }
    InvParams p;
    memset(&p, 0, sizeof(p));
    p.n = n_bus; p.n_load = n_load; p.B = B; p.By = By; p.Bl = Bl; p.Bv = Bv;
    p.y_base = (const double2*)y_base; p.s_load = (const double2*)s_load; p.v = (const double2*)v;
    p.y_load = (double2*)y_load; p.y_diag = (double2*)y_diag; p.z0 = (double2*)z0;
    for (int l = 0; l < n_load; ++l) {
        if (load_bus[l] < 0 || load_bus[l] >= n_bus) { set_error("dps_dynamic_inverse: load bus out of range"); return 1; }
        p.load_bus[l] = (unsigned char)load_bus[l];
    }
    const size_t smem = sizeof(double2) * ((size_t)n_bus * 2 * n_bus + n_bus) + sizeof(double) * INV_THREADS + sizeof(int) * INV_THREADS;
    cudaError_t e = cudaFuncSetAttribute(dps_dyn_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("dps_dynamic_inverse: %s", cudaGetErrorString(e)); return 3; }
    dps_dyn_inverse_kernel<<<(unsigned)B, INV_THREADS, smem, (cudaStream_t)stream>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("dps_dynamic_inverse: %s", cudaGetErrorString(e)); return 3; }
    return 0;
}

int dps_model_init(int32_t n_gen, int32_t n_bus, int64_t B, const int32_t* gen_bus, const int32_t* ctrl, const double* v,
                   int64_t Bv, const double* s, int64_t Bs, const double* raw_u, const int32_t* raw_vary, const double* raw_v,
                   double* x0, double* slots, double* inj, void* stream) {
    if (n_gen < 1 || n_gen > MAX_GW) { set_error("dps_model_init: %d generators (1..%d)", n_gen, MAX_GW); return 2; }
    if (n_bus < 1 || n_bus > 32767) { set_error("dps_model_init: %d buses (1..32767)", n_bus); return 2; }
    if (B < 1 || !gen_bus || !ctrl || !v || !s || !raw_u || !raw_vary || !x0 || !slots || !inj) { set_error("dps_model_init: bad argument"); return 1; }
    if ((Bv != 1 && Bv != B) || (Bs != 1 && Bs != B)) { set_error("dps_model_init: batch sizes must be 1 or B"); return 1; }
    static thread_local InitParams p;
    memset(&p, 0, sizeof(p));
    p.n_gen = n_gen; p.n_bus = n_bus; p.B = B; p.Bv = Bv; p.Bs = Bs;
    p.v = (const double2*)v; p.s = (const double2*)s; p.raw_u = raw_u; p.raw_v = raw_v; p.x0 = x0; p.slots = slots; p.inj = inj;
    for (int g = 0; g < n_gen; ++g) {
        if (gen_bus[g] < 0 || gen_bus[g] >= n_bus) { set_error("dps_model_init: generator bus out of range"); return 1; }
        p.gen_bus[g] = (short)gen_bus[g];
        p.ctrl[g] = (unsigned char)ctrl[g];
        for (int k = 0; k < DPS_N_RAW; ++k) {
            const int vi = raw_vary[g * DPS_N_RAW + k];
            if (vi >= 0 && !raw_v) { set_error("dps_model_init: per-trajectory parameters requested but raw_v is null"); return 1; }
            if (vi > 32767) { set_error("dps_model_init: too many per-trajectory parameter rows"); return 2; }
            p.raw_vary[g][k] = (short)vi;
        }
    }
    const int threads = 128;
    dim3 grid((unsigned)((B + threads - 1) / threads), (unsigned)n_gen);
    dps_model_init_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("dps_model_init: %s", cudaGetErrorString(e)); return 3; }
    return 0;
}

}  // extern "C"
