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

// ----------------------------------------------------------------------------------------------------------------------
// host emulation of the adjoint (csrc/dps_adjoint.cuh + the sweep structure of csrc/dps_backward.cu)
// ----------------------------------------------------------------------------------------------------------------------
#include "../../diffpssi_b200/csrc/dps_adjoint.cuh"

extern "C" int emul_backward(int n_gen, int solver, int n_steps, double dt, double fn, const int* ctrl, int n_obs,
                             int z_rows, int z_cols, int n_cfg, const uint8_t* step_cfg, int z_per_sim,
                             const int* slot_vary, int n_rec, const int* rec_chan, long long P, const double* x0,
                             const double* upar, const double* vpar, const double* zsub, const double* grad_rec,
                             double* g_x0, double* gslot_all /* [n_gen][DPS_N_SLOTS][P] */,
                             double* g_zsub /* [n_cfg][z_rows][z_cols][2][P] accumulated, or null */) {
    const int G = n_gen, NS = DPS_N_STATES;
    const int n_stages = solver == DPS_EULER ? 1 : (solver == DPS_HEUN ? 2 : 4);
    const long long zb = z_per_sim ? P : 1;
    double cst[4] = {0, 0, 0, 0}, bst[4] = {0, 0, 0, 0};
    if (solver == DPS_EULER) { bst[0] = 1.0; }
    else if (solver == DPS_HEUN) { cst[1] = 1.0; bst[0] = bst[1] = 0.5; }
    else { cst[1] = 0.5; cst[2] = 0.5; cst[3] = 1.0; bst[0] = 1.0 / 6.0; bst[1] = 1.0 / 3.0; bst[2] = 1.0 / 3.0; bst[3] = 1.0 / 6.0; }
    std::vector<short> sv(G * DPS_N_SLOTS);
    for (int i = 0; i < G * DPS_N_SLOTS; ++i) sv[i] = (short)slot_vary[i];
    (void)n_cfg;
    for (long long s = 0; s < P; ++s) {
        const long long zs = z_per_sim ? s : 0;
        auto z_at = [&](int cfg, int row, int col, double& zr, double& zi) {
            const double* e = zsub + ((((long long)cfg * z_rows + row) * z_cols + col) * 2) * zb + zs;
            zr = e[0];
            zi = e[zb];
        };
        auto par_of = [&](int g) { return ParamReader<true>{upar + g * DPS_N_SLOTS, sv.data() + g * DPS_N_SLOTS, vpar, P, s}; };
        // all stage inputs of one step: X[i][g][k], sd/cd/v per (i, g)
        auto eval_all = [&](const std::vector<double>& X, int cfg, std::vector<double>& sd, std::vector<double>& cd,
                            std::vector<double>& vre, std::vector<double>& vim, std::vector<double>& f) {
            std::vector<double> ire(G), iim(G);
            for (int g = 0; g < G; ++g) {
                auto par = par_of(g);
                sincos(X[g * NS + DPS_X_DELTA], &sd[g], &cd[g]);
                injection(&X[g * NS], par, sd[g], cd[g], ire[g], iim[g]);
            }
            for (int g = 0; g < G; ++g) {
                auto par = par_of(g);
                double v_re = 0, v_im = 0;
                for (int j = 0; j < G; ++j) {
                    double zr, zi;
                    z_at(cfg, g, j, zr, zi);
                    v_re += zr * ire[j] - zi * iim[j];
                    v_im += zr * iim[j] + zi * ire[j];
                }
                vre[g] = v_re;
                vim[g] = v_im;
                Algebraics alg;
                stage_derivatives<true>(&X[g * NS], par, ctrl[g], fn, sd[g], cd[g], v_re, v_im, &f[g * NS], alg);
            }
        };
        // forward pass storing the state at the start of every step (the emulation keeps the whole trajectory)
        std::vector<std::vector<double>> traj(n_steps, std::vector<double>(G * NS));
        std::vector<double> x(G * NS), f(G * NS), acc(G * NS), xs(G * NS), sd(G), cd(G), vre(G), vim(G);
        for (int g = 0; g < G; ++g)
            for (int k = 0; k < NS; ++k) x[g * NS + k] = x0[((long long)g * NS + k) * P + s];
        for (int step = 0; step < n_steps; ++step) {
            traj[step] = x;
            xs = x;
            for (int st = 0; st < n_stages; ++st) {
                eval_all(x, step_cfg[step], sd, cd, vre, vim, f);
                for (int i = 0; i < G * NS; ++i) {
                    if (solver == DPS_EULER) x[i] = xs[i] + f[i] * dt;
                    else if (solver == DPS_HEUN) {
                        if (st == 0) { acc[i] = f[i]; x[i] = xs[i] + f[i] * dt; }
                        else x[i] = xs[i] + ((acc[i] + f[i]) * 0.5) * dt;
                    } else {
                        if (st == 0) { acc[i] = f[i]; x[i] = xs[i] + (f[i] * dt) * 0.5; }
                        else if (st == 1) { acc[i] += 2.0 * f[i]; x[i] = xs[i] + (f[i] * dt) * 0.5; }
                        else if (st == 2) { acc[i] += 2.0 * f[i]; x[i] = xs[i] + f[i] * dt; }
                        else x[i] = xs[i] + ((acc[i] + f[i]) * dt) / 6.0;
                    }
                }
            }
        }
        std::vector<double> lam(G * NS, 0.0), gp(G * DPS_N_SLOTS, 0.0);
        for (int step = n_steps - 1; step >= 0; --step) {
            const int cfg = step_cfg[step];
            for (int c = 0; c < n_rec; ++c)
                if (rec_chan[3 * c] == DPS_CH_GEN && rec_chan[3 * c + 2] < NS)
                    lam[rec_chan[3 * c + 1] * NS + rec_chan[3 * c + 2]] += grad_rec[((long long)c * n_steps + step) * P + s];
            std::vector<std::vector<double>> X(n_stages, std::vector<double>(G * NS)), SD(n_stages, std::vector<double>(G)),
                CD(n_stages, std::vector<double>(G)), VR(n_stages, std::vector<double>(G)), VI(n_stages, std::vector<double>(G));
            X[0] = traj[step];
            for (int i = 0; i < n_stages; ++i) {
                eval_all(X[i], cfg, SD[i], CD[i], VR[i], VI[i], f);
                if (i + 1 < n_stages)
                    for (int k = 0; k < G * NS; ++k) X[i + 1][k] = X[0][k] + cst[i + 1] * (f[k] * dt);
            }
            std::vector<double> sum_dx(G * NS, 0.0), dxn(G * NS, 0.0);
            for (int i = n_stages - 1; i >= 0; --i) {
                std::vector<double> dk(G * NS), dx(G * NS, 0.0), dvr(G), dvi(G), dsd(G), dcd(G);
                const double bi = bst[i] * dt, cn = (i + 1 < n_stages) ? cst[i + 1] * dt : 0.0;
                for (int k = 0; k < G * NS; ++k) dk[k] = bi * lam[k] + cn * dxn[k];
                for (int g = 0; g < G; ++g) {
                    auto par = par_of(g);
                    AlgCot ac = {0, 0, 0, 0, 0, 0, 0};
                    if (i == n_stages - 1) {
                        double* av[7] = {&ac.p_e, &ac.i_d, &ac.i_q, &ac.v_re, &ac.v_im, &ac.e_fd, &ac.p_m};
                        for (int c = 0; c < n_rec; ++c)
                            if (rec_chan[3 * c] == DPS_CH_GEN && rec_chan[3 * c + 1] == g && rec_chan[3 * c + 2] >= NS)
                                *av[rec_chan[3 * c + 2] - NS] += grad_rec[((long long)c * n_steps + step) * P + s];
                    }
                    double* gpg = &gp[g * DPS_N_SLOTS];
                    stage_vjp<true>(&X[i][g * NS], par, ctrl[g], fn, SD[i][g], CD[i][g], VR[i][g], VI[i][g], &dk[g * NS], ac,
                                    &dx[g * NS], gpg, dsd[g], dcd[g], dvr[g], dvi[g]);
                }
                for (int j = 0; j < G; ++j) {
                    auto par = par_of(j);
                    double di_re = 0, di_im = 0;
                    double own_re = 0, own_im = 0;
                    if (g_zsub) injection(&X[i][j * NS], par, SD[i][j], CD[i][j], own_re, own_im);
                    auto gz_at = [&](int r, int c) { return g_zsub + ((((long long)cfg * z_rows + r) * z_cols + j) * 2 + c) * P + s; };
                    for (int r = 0; r < G; ++r) {
                        double zr, zi;
                        z_at(cfg, r, j, zr, zi);
                        di_re += zr * dvr[r] + zi * dvi[r];
                        di_im += -zi * dvr[r] + zr * dvi[r];
                        if (g_zsub) {
                            *gz_at(r, 0) += dvr[r] * own_re + dvi[r] * own_im;
                            *gz_at(r, 1) += -dvr[r] * own_im + dvi[r] * own_re;
                        }
                    }
                    if (i == n_stages - 1) {
                        for (int c = 0; c < n_rec; ++c) {
                            const int kind = rec_chan[3 * c];
                            if (kind == DPS_CH_GEN) continue;
                            const int r = rec_chan[3 * c + 1];
                            const double a = grad_rec[((long long)c * n_steps + step) * P + s];
                            double zr, zi;
                            z_at(cfg, r, j, zr, zi);
                            if (kind == DPS_CH_BUS_VRE) { di_re += zr * a; di_im += -zi * a; }
                            else { di_re += zi * a; di_im += zr * a; }
                            if (g_zsub) {
                                const double ar = (kind == DPS_CH_BUS_VRE) ? a : 0.0, ai = (kind == DPS_CH_BUS_VRE) ? 0.0 : a;
                                *gz_at(r, 0) += ar * own_re + ai * own_im;
                                *gz_at(r, 1) += -ar * own_im + ai * own_re;
                            }
                        }
                    }
                    double* gpj = &gp[j * DPS_N_SLOTS];
                    injection_vjp(&X[i][j * NS], par, SD[i][j], CD[i][j], di_re, di_im, &dx[j * NS], gpj, dsd[j], dcd[j]);
                    dx[j * NS + DPS_X_DELTA] += dsd[j] * CD[i][j] - dcd[j] * SD[i][j];
                }
                dxn = dx;
                for (int k = 0; k < G * NS; ++k) sum_dx[k] += dx[k];
            }
            for (int k = 0; k < G * NS; ++k) lam[k] += sum_dx[k];
        }
        for (int g = 0; g < G; ++g) {
            for (int k = 0; k < NS; ++k) g_x0[((long long)g * NS + k) * P + s] = lam[g * NS + k];
            for (int k = 0; k < DPS_N_SLOTS; ++k) gslot_all[((long long)g * DPS_N_SLOTS + k) * P + s] = gp[g * DPS_N_SLOTS + k];
        }
    }
    return 0;
}
