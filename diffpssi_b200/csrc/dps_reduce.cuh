// dps_reduce.cuh -- Sherman-Morrison fault re-reduction for ONE trajectory (host/device), see dps_network.cu.
#pragma once
#include "dps_internal.h"

#ifndef DPS_HD
#ifdef __CUDACC__
#define DPS_HD __host__ __device__ __forceinline__
#else
#define DPS_HD inline
#endif
#endif

namespace dps {

DPS_HD void network_reduce_one(const ReduceParams& p, long long b) {
    const long long B = p.Bz;
    const int rows = p.n_obs + p.m_max, cols = p.n_gen + p.m_max;
    for (int cfg = 0; cfg < p.n_cfg; ++cfg) {
        double* W = p.zsub + (long long)cfg * rows * cols * 2 * B + b;
        const int m = p.cfg_n[cfg];
        auto w_re = [&](int r, int c) -> double& { return W[((long long)(r * cols + c) * 2) * B]; };
        auto w_im = [&](int r, int c) -> double& { return W[((long long)(r * cols + c) * 2 + 1) * B]; };
        // gather rows {observed, faulted} x columns {generators, faulted}
        for (int r = 0; r < rows; ++r) {
            const int rb = (r < p.n_obs) ? p.obs_bus[r] : ((r - p.n_obs) < m ? p.cfg_bus[cfg][r - p.n_obs] : -1);
            for (int c = 0; c < cols; ++c) {
                const int cbus = (c < p.n_gen) ? p.gen_bus[c] : ((c - p.n_gen) < m ? p.cfg_bus[cfg][c - p.n_gen] : -1);
                double re = 0.0, im = 0.0;
                if (rb >= 0 && cbus >= 0) {
                    const long long o = ((long long)rb * p.n_bus + cbus) * B + b;
                    re = p.z_re[o];
                    im = p.z_im[o];
                }
                w_re(r, c) = re;
                w_im(r, c) = im;
            }
        }
        // one Sherman-Morrison update per faulted bus (the diagonal of Y is SET to y_fault: simulator.py:527)
        for (int k = 0; k < m; ++k) {
            const int fb = p.cfg_bus[cfg][k];
            const int rk = p.n_obs + k, ck = p.n_gen + k;
            const double c_re = p.y_fault - p.yd_re[(long long)fb * B + b];
            const double c_im = -p.yd_im[(long long)fb * B + b];
            const double zbb_re = w_re(rk, ck), zbb_im = w_im(rk, ck);
            // den = 1 + c * Z_bb ; f = c / den
            const double den_re = 1.0 + (c_re * zbb_re - c_im * zbb_im);
            const double den_im = c_re * zbb_im + c_im * zbb_re;
            const double dn = den_re * den_re + den_im * den_im;
            const double f_re = (c_re * den_re + c_im * den_im) / dn;
            const double f_im = (c_im * den_re - c_re * den_im) / dn;
            // rows other than the faulted one first (they need the OLD faulted row), then the faulted row itself
            for (int pass = 0; pass < 2; ++pass) {
                for (int r = 0; r < rows; ++r) {
                    if ((r == rk) != (pass == 1)) continue;
                    const double a_re = w_re(r, ck), a_im = w_im(r, ck);           // Z[r][b] (old)
                    const double t_re = f_re * a_re - f_im * a_im;
                    const double t_im = f_re * a_im + f_im * a_re;                  // f * Z[r][b]
                    for (int c = 0; c < cols; ++c) {
                        const double b_re = w_re(rk, c), b_im = w_im(rk, c);       // Z[b][c] (old: for r==rk it is this entry)
                        w_re(r, c) -= t_re * b_re - t_im * b_im;
                        w_im(r, c) -= t_re * b_im + t_im * b_re;
                    }
                }
            }
        }
    }
}

}  // namespace dps
