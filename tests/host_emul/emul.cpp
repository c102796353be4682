// TEST INFRASTRUCTURE -- host build of the *device* model code for CPU-only pre-validation.
//
// Compiles csrc/dps_model.cuh and csrc/dps_reduce.cuh (the very headers the CUDA kernels include) with g++ and
// drives them with a scalar loop that mirrors the indexing of dps_step_kernel (csrc/dps_forward.cu): same packed
// buffers (x0, upar, vpar, zsub, rec), same update formulas, same recorder layout.  It lets the `-m "not gpu"` suite
// check the host packing (planning.py), the slot derivation (prologue.py) and the shared model equations against the
// oracle and the golden vectors before any GPU time is spent.  It is NOT a fallback: nothing in the product package
// can load it (it lives under tests/ and engine.py only ever dlopens libdps_b200.so).
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <vector>

#define __device__
#define __forceinline__ inline
#define __ldg(p) (*(p))
#include "../../diffpssi_b200/csrc/dps_model.cuh"
#include "../../diffpssi_b200/csrc/dps_reduce.cuh"

using namespace dps;

extern "C" {

int emul_network_reduce(long long Bz, int n_bus, const double* z_re, const double* z_im, const double* yd_re,
                        const double* yd_im, int n_obs, const int* obs_bus, int n_gen, const int* gen_bus, int n_cfg,
                        const int* cfg_ptr, const int* cfg_bus, double y_fault, double* zsub) {
    ReduceParams p;
    memset(&p, 0, sizeof(p));
    p.Bz = Bz; p.n_bus = n_bus; p.n_obs = n_obs; p.n_gen = n_gen; p.n_cfg = n_cfg; p.y_fault = y_fault;
    p.z_re = z_re; p.z_im = z_im; p.yd_re = yd_re; p.yd_im = yd_im; p.zsub = zsub;
    for (int i = 0; i < n_obs; ++i) p.obs_bus[i] = obs_bus[i];
    for (int i = 0; i < n_gen; ++i) p.gen_bus[i] = gen_bus[i];
    int m_max = 0;
    for (int c = 0; c < n_cfg; ++c) {
        const int m = cfg_ptr[c + 1] - cfg_ptr[c];
        p.cfg_n[c] = m;
        for (int k = 0; k < m; ++k) p.cfg_bus[c][k] = cfg_bus[cfg_ptr[c] + k];
        if (m > m_max) m_max = m;
    }
    p.m_max = m_max;
    for (long long b = 0; b < Bz; ++b) network_reduce_one(p, b);
    return 0;
}

// rec_chan: [n_rec][3] (kind, index, field) exactly as passed to dps_plan_create
int emul_forward(int n_gen, int solver, int n_steps, double dt, double fn, const int* ctrl, int n_obs, int z_rows,
                 int z_cols, int n_cfg, const uint8_t* step_cfg, int z_per_sim, const int* slot_vary, int n_rec,
                 const int* rec_chan, long long P, const double* x0, const double* upar, const double* vpar,
                 const double* zsub, double* rec, double* x_final) {
    const int G = n_gen;
    const int n_stages = solver == DPS_EULER ? 1 : (solver == DPS_HEUN ? 2 : 4);
    const long long zb = z_per_sim ? P : 1;
    std::vector<short> sv(G * DPS_N_SLOTS);
    for (int i = 0; i < G * DPS_N_SLOTS; ++i) sv[i] = (short)slot_vary[i];
    bool any_ctrl = false;
    for (int g = 0; g < G; ++g) any_ctrl |= ctrl[g] != 0;
    (void)n_cfg;
    for (long long s = 0; s < P; ++s) {
        const long long zs = z_per_sim ? s : 0;
        std::vector<double> x(G * DPS_N_STATES), xs(G * DPS_N_STATES), acc(G * DPS_N_STATES), f(G * DPS_N_STATES);
        std::vector<double> ire(G), iim(G), sd(G), cd(G);
        std::vector<Algebraics> alg(G);
        for (int g = 0; g < G; ++g)
            for (int k = 0; k < DPS_N_STATES; ++k) x[g * DPS_N_STATES + k] = x0[((long long)g * DPS_N_STATES + k) * P + s];
        for (int step = 0; step < n_steps; ++step) {
            const int cfg = step_cfg[step];
            xs = x;
            std::vector<double> busv(2 * (n_obs > G ? n_obs - G : 0));
            for (int st = 0; st < n_stages; ++st) {
                for (int g = 0; g < G; ++g) {
                    ParamReader<true> par{upar + g * DPS_N_SLOTS, sv.data() + g * DPS_N_SLOTS, vpar, P, s};
                    sincos(x[g * DPS_N_STATES + DPS_X_DELTA], &sd[g], &cd[g]);
                    injection(&x[g * DPS_N_STATES], par, sd[g], cd[g], ire[g], iim[g]);
                }
                for (int g = 0; g < G; ++g) {
                    ParamReader<true> par{upar + g * DPS_N_SLOTS, sv.data() + g * DPS_N_SLOTS, vpar, P, s};
                    const double* base = zsub + (((long long)cfg * z_rows + g) * z_cols) * 2 * zb + zs;
                    double v_re = 0.0, v_im = 0.0;
                    for (int j = 0; j < G; ++j) {
                        const double zr = base[(long long)(2 * j) * zb], zi = base[(long long)(2 * j + 1) * zb];
                        v_re += zr * ire[j] - zi * iim[j];
                        v_im += zr * iim[j] + zi * ire[j];
                    }
                    if (any_ctrl)
                        stage_derivatives<true>(&x[g * DPS_N_STATES], par, ctrl[g], fn, sd[g], cd[g], v_re, v_im,
                                                &f[g * DPS_N_STATES], alg[g]);
                    else
                        stage_derivatives<false>(&x[g * DPS_N_STATES], par, 0, fn, sd[g], cd[g], v_re, v_im,
                                                 &f[g * DPS_N_STATES], alg[g]);
                }
                if (st == n_stages - 1) {
                    for (int r = G; r < n_obs; ++r) {
                        const double* base = zsub + (((long long)cfg * z_rows + r) * z_cols) * 2 * zb + zs;
                        double br = 0.0, bi = 0.0;
                        for (int j = 0; j < G; ++j) {
                            const double zr = base[(long long)(2 * j) * zb], zi = base[(long long)(2 * j + 1) * zb];
                            br += zr * ire[j] - zi * iim[j];
                            bi += zr * iim[j] + zi * ire[j];
                        }
                        busv[2 * (r - G)] = br;
                        busv[2 * (r - G) + 1] = bi;
                    }
                }
                for (int i = 0; i < G * DPS_N_STATES; ++i) {
                    if (solver == DPS_EULER) {
                        x[i] = xs[i] + f[i] * dt;
                    } else if (solver == DPS_HEUN) {
                        if (st == 0) { acc[i] = f[i]; x[i] = xs[i] + f[i] * dt; }
                        else x[i] = xs[i] + ((acc[i] + f[i]) / 2.0) * dt;
                    } else {
                        if (st == 0) { acc[i] = f[i]; x[i] = xs[i] + (f[i] * dt) / 2.0; }
                        else if (st == 1) { acc[i] = acc[i] + 2.0 * f[i]; x[i] = xs[i] + (f[i] * dt) / 2.0; }
                        else if (st == 2) { acc[i] = acc[i] + 2.0 * f[i]; x[i] = xs[i] + f[i] * dt; }
                        else x[i] = xs[i] + ((acc[i] + f[i]) * dt) / 6.0;
                    }
                }
            }
            for (int c = 0; c < n_rec; ++c) {
                const int kind = rec_chan[3 * c], idx = rec_chan[3 * c + 1], field = rec_chan[3 * c + 2];
                double v;
                if (kind == DPS_CH_GEN) {
                    if (field < DPS_N_STATES) v = x[idx * DPS_N_STATES + field];
                    else {
                        const Algebraics& a = alg[idx];
                        const double av[7] = {a.p_e, a.i_d, a.i_q, a.v_re, a.v_im, a.e_fd, a.p_m};
                        v = av[field - DPS_N_STATES];
                    }
                } else {
                    v = busv[2 * (idx - G) + (kind == DPS_CH_BUS_VIM)];
                }
                rec[((long long)c * n_steps + step) * P + s] = v;
            }
        }
        if (x_final)
            for (int g = 0; g < G; ++g)
                for (int k = 0; k < DPS_N_STATES; ++k) x_final[((long long)g * DPS_N_STATES + k) * P + s] = x[g * DPS_N_STATES + k];
    }
    return 0;
}

}  // extern "C"
