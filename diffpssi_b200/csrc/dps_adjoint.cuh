// dps_adjoint.cuh -- hand-written vector-Jacobian products of the stage evaluation in dps_model.cuh.
//
// This is what torch autograd derives automatically for the reference when ps_optimization.py:212 calls
// loss.sum().backward(): the transpose of every op of update_internal_vars / differential / the controller blocks
// (synchronous_machine.py:138-277, blocks.py, exciters.py, governors.py, stabilizers.py).  Gate conventions follow
// torch:
//   * torch.where(cond, 0, dx)      (blocks.py:145-149): no gradient through a frozen PT1 derivative;
//   * minimum(maximum(x, lo), hi)   (blocks.py:181, 241): gradient 1 inside, 0 outside, 1/2 on an exact tie;
//   * abs(v) of a complex voltage   (synchronous_machine.py:151): d|v| = (v_re dv_re + v_im dv_im)/|v|, 0 at v = 0.
#pragma once
#include "dps_model.cuh"

namespace dps {

// d clamp(x, lo, hi) / dx with torch's tie rule (maximum then minimum, each splitting 1/2 : 1/2 on equality)
__device__ __forceinline__ double clamp_dx(double x, double lo, double hi) {
    const double g1 = (x > lo) ? 1.0 : ((x == lo) ? 0.5 : 0.0);        // maximum(x, lo) wrt x
    const double m = fmax(x, lo);
    const double g2 = (m < hi) ? 1.0 : ((m == hi) ? 0.5 : 0.0);        // minimum(m, hi) wrt m
    return g1 * g2;
}
__device__ __forceinline__ double clamp_dlo(double x, double lo, double hi) {
    const double g1 = (lo > x) ? 1.0 : ((x == lo) ? 0.5 : 0.0);
    const double m = fmax(x, lo);
    const double g2 = (m < hi) ? 1.0 : ((m == hi) ? 0.5 : 0.0);
    return g1 * g2;
}
__device__ __forceinline__ double clamp_dhi(double x, double lo, double hi) {
    const double m = fmax(x, lo);
    return (hi < m) ? 1.0 : ((m == hi) ? 0.5 : 0.0);
}
__device__ __forceinline__ bool windup_stop(double state, double dx, double lo, double hi) {
    return ((state <= lo) && (dx < 0.0)) || ((state >= hi) && (dx > 0.0));
}

struct AlgCot {                 // cotangents of the recorded algebraics (non-zero only on the last stage of a step)
    double p_e, i_d, i_q, v_re, v_im, e_fd, p_m;
};

// Transpose of stage_derivatives().  Inputs: the stage's forward inputs (x, sd, cd, v), the cotangent df[13] of the
// derivatives and the cotangent `ac` of the algebraics.  Accumulates into dx[13], gp[DPS_N_SLOTS] and returns the
// cotangents of (sin d, cos d) and of the bus voltage.
template <bool CTRL, class PR, class GP>
__device__ __forceinline__ void stage_vjp(const double* x, const PR& par, int ctrl, double fn, double sd, double cd,
                                          double v_re, double v_im, const double* df, const AlgCot& ac,
                                          double* dx, GP& gp, double& dsd, double& dcd, double& dv_re, double& dv_im) {
    const double omega = x[DPS_X_OMEGA];
    const double eqst = x[DPS_X_EQST], edst = x[DPS_X_EDST], eqt = x[DPS_X_EQT], edt = x[DPS_X_EDT];
    // ---- forward recompute of the algebraic chain --------------------------------------------------------------------
    const double e_re = eqst * cd + edst * sd;
    const double e_im = eqst * sd - edst * cd;
    const double ixd = par(DPS_S_INV_XDST);
    const double i_re = (e_im - v_im) * ixd;
    const double i_im = -((e_re - v_re) * ixd);
    const double p_e = v_re * i_re + v_im * i_im;
    const double i_d = i_re * sd - i_im * cd;
    const double i_q = i_re * cd + i_im * sd;

    double d_efd = ac.e_fd, d_pm = ac.p_m, d_pe = ac.p_e, d_id = ac.i_d, d_iq = ac.i_q;
    double d_omega = 0.0;
    dv_re = ac.v_re;
    dv_im = ac.v_im;
    dsd = 0.0;
    dcd = 0.0;

    // values needed from the controller forward pass
    double e_fd = par(DPS_S_EFD0), p_m = par(DPS_S_PM0);
    double w = 0.0, o1 = 0.0, o2 = 0.0, v_pss = 0.0, v_abs = 0.0, u_a = 0.0, o_a = 0.0, u_g = 0.0, gs = 0.0, o_g = 0.0;
    const bool has_pss = CTRL && (ctrl & DPS_CTRL_PSS), has_avr = CTRL && (ctrl & DPS_CTRL_AVR),
               has_gov = CTRL && (ctrl & DPS_CTRL_GOV);
    const bool fed = has_pss && has_avr;
    if (CTRL) {
        if (has_pss) {
            w = par(DPS_S_S_INV_TW) * (par(DPS_S_S_KW) * (fed ? omega : 0.0) - x[DPS_X_PSS_W]);
            o1 = par(DPS_S_S_R1) * (fed ? w : 0.0) + par(DPS_S_S_OMR1) * x[DPS_X_PSS_LL1];
            o2 = par(DPS_S_S_R2) * (fed ? o1 : 0.0) + par(DPS_S_S_OMR2) * x[DPS_X_PSS_LL2];
            v_pss = clampd(o2, -par(DPS_S_S_HLIM), par(DPS_S_S_HLIM));
        }
        if (has_avr) {
            v_abs = dps_sqrt(v_re * v_re + v_im * v_im);
            const double sig = has_pss ? (v_abs - v_pss) : v_abs;
            u_a = (par(DPS_S_A_VSET) - sig) + par(DPS_S_A_BIAS);
            o_a = par(DPS_S_A_R) * u_a + par(DPS_S_A_OMR) * x[DPS_X_AVR_LL];
            e_fd = clampd(x[DPS_X_AVR_PT1], par(DPS_S_A_EMIN), par(DPS_S_A_EMAX));
        }
        if (has_gov) {
            u_g = par(DPS_S_G_PREF) - omega;
            gs = clampd(x[DPS_X_GOV_PT1], par(DPS_S_G_VMIN), par(DPS_S_G_VMAX));
            o_g = par(DPS_S_G_R) * gs + par(DPS_S_G_OMR) * x[DPS_X_GOV_LL];
            p_m = o_g - par(DPS_S_G_DT) * omega;
        }
    }
    const double inv1w = dps_rcp(1.0 + omega);
    const double t_m = p_m * inv1w;

    // ---- machine equations, synchronous_machine.py:158-164 -----------------------------------------------------------
    {
        const double a = df[DPS_X_OMEGA];
        gp[DPS_S_INV_2H] += a * ((t_m - p_e) - par(DPS_S_D) * omega);
        const double t = a * par(DPS_S_INV_2H);
        const double d_tm = t;
        d_pe -= t;
        gp[DPS_S_D] -= t * omega;
        d_omega -= t * par(DPS_S_D);
        d_pm += d_tm * inv1w;                           // t_m = p_m/(1+omega)
        d_omega -= d_tm * t_m * inv1w;
    }
    d_omega += df[DPS_X_DELTA] * (((2.0) * 3.141592653589793) * fn);
    {
        const double a = df[DPS_X_EQT];
        gp[DPS_S_INV_TD0T] += a * ((e_fd - eqt) - i_d * par(DPS_S_XD_XDT));
        const double t = a * par(DPS_S_INV_TD0T);
        d_efd += t;
        dx[DPS_X_EQT] -= t;
        d_id -= t * par(DPS_S_XD_XDT);
        gp[DPS_S_XD_XDT] -= t * i_d;
    }
    {
        const double a = df[DPS_X_EDT];
        gp[DPS_S_INV_TQ0T] += a * (-edt + i_q * par(DPS_S_XQ_XQT));
        const double t = a * par(DPS_S_INV_TQ0T);
        dx[DPS_X_EDT] -= t;
        d_iq += t * par(DPS_S_XQ_XQT);
        gp[DPS_S_XQ_XQT] += t * i_q;
    }
    {
        const double a = df[DPS_X_EQST];
        gp[DPS_S_INV_TD0ST] += a * ((eqt - eqst) - i_d * par(DPS_S_XDT_XDST));
        const double t = a * par(DPS_S_INV_TD0ST);
        dx[DPS_X_EQT] += t;
        dx[DPS_X_EQST] -= t;
        d_id -= t * par(DPS_S_XDT_XDST);
        gp[DPS_S_XDT_XDST] -= t * i_d;
    }
    {
        const double a = df[DPS_X_EDST];
        gp[DPS_S_INV_TQ0ST] += a * ((edt - edst) + i_q * par(DPS_S_XQT_XQST));
        const double t = a * par(DPS_S_INV_TQ0ST);
        dx[DPS_X_EDT] += t;
        dx[DPS_X_EDST] -= t;
        d_iq += t * par(DPS_S_XQT_XQST);
        gp[DPS_S_XQT_XQST] += t * i_q;
    }

    // ---- controllers -------------------------------------------------------------------------------------------------
    double d_vpss = 0.0;
    if (CTRL && has_gov) {
        const double xg = x[DPS_X_GOV_PT1], xll = x[DPS_X_GOV_LL];
        const double vmin = par(DPS_S_G_VMIN), vmax = par(DPS_S_G_VMAX);
        double d_o = d_pm;                               // p_m = o - D_t*omega
        gp[DPS_S_G_DT] -= d_pm * omega;
        d_omega -= d_pm * par(DPS_S_G_DT);
        double d_gs = 0.0, d_u = 0.0;
        {   // f9 = inv_t3*(gs - xll)
            const double a = df[DPS_X_GOV_LL];
            gp[DPS_S_G_INV_T3] += a * (gs - xll);
            const double t = a * par(DPS_S_G_INV_T3);
            d_gs += t;
            dx[DPS_X_GOV_LL] -= t;
        }
        {   // f8 = gate(inv_t1*(gain*u - xg))
            const double raw = par(DPS_S_G_INV_T1) * (par(DPS_S_G_GAIN) * u_g - xg);
            const double a = windup_stop(xg, raw, vmin, vmax) ? 0.0 : df[DPS_X_GOV_PT1];
            gp[DPS_S_G_INV_T1] += a * (par(DPS_S_G_GAIN) * u_g - xg);
            const double t = a * par(DPS_S_G_INV_T1);
            gp[DPS_S_G_GAIN] += t * u_g;
            d_u += t * par(DPS_S_G_GAIN);
            dx[DPS_X_GOV_PT1] -= t;
        }
        gp[DPS_S_G_R] += d_o * gs;                       // o = r*gs + omr*xll
        d_gs += d_o * par(DPS_S_G_R);
        gp[DPS_S_G_OMR] += d_o * xll;
        dx[DPS_X_GOV_LL] += d_o * par(DPS_S_G_OMR);
        dx[DPS_X_GOV_PT1] += d_gs * clamp_dx(xg, vmin, vmax);
        gp[DPS_S_G_VMIN] += d_gs * clamp_dlo(xg, vmin, vmax);
        gp[DPS_S_G_VMAX] += d_gs * clamp_dhi(xg, vmin, vmax);
        gp[DPS_S_G_PREF] += d_u;                         // u = p_ref - omega
        d_omega -= d_u;
    } else {
        gp[DPS_S_PM0] += d_pm;
    }
    if (CTRL && has_avr) {
        const double xll = x[DPS_X_AVR_LL], xe = x[DPS_X_AVR_PT1];
        const double emin = par(DPS_S_A_EMIN), emax = par(DPS_S_A_EMAX);
        double d_u = 0.0, d_o = 0.0;
        {   // f7 = gate(inv_te*(K*o - xe))
            const double raw = par(DPS_S_A_INV_TE) * (par(DPS_S_A_K) * o_a - xe);
            const double a = windup_stop(xe, raw, emin, emax) ? 0.0 : df[DPS_X_AVR_PT1];
            gp[DPS_S_A_INV_TE] += a * (par(DPS_S_A_K) * o_a - xe);
            const double t = a * par(DPS_S_A_INV_TE);
            gp[DPS_S_A_K] += t * o_a;
            d_o += t * par(DPS_S_A_K);
            dx[DPS_X_AVR_PT1] -= t;
        }
        {   // f6 = inv_tb*(u - xll)
            const double a = df[DPS_X_AVR_LL];
            gp[DPS_S_A_INV_TB] += a * (u_a - xll);
            const double t = a * par(DPS_S_A_INV_TB);
            d_u += t;
            dx[DPS_X_AVR_LL] -= t;
        }
        dx[DPS_X_AVR_PT1] += d_efd * clamp_dx(xe, emin, emax);          // e_fd = clamp(xe)
        gp[DPS_S_A_EMIN] += d_efd * clamp_dlo(xe, emin, emax);
        gp[DPS_S_A_EMAX] += d_efd * clamp_dhi(xe, emin, emax);
        gp[DPS_S_A_R] += d_o * u_a;                                     // o = r*u + omr*xll
        d_u += d_o * par(DPS_S_A_R);
        gp[DPS_S_A_OMR] += d_o * xll;
        dx[DPS_X_AVR_LL] += d_o * par(DPS_S_A_OMR);
        gp[DPS_S_A_VSET] += d_u;                                        // u = (vset - sig) + bias
        gp[DPS_S_A_BIAS] += d_u;
        const double d_sig = -d_u;
        if (has_pss) d_vpss = -d_sig;                                   // sig = |v| - v_pss
        if (v_abs > 0.0) {
            const double t = d_sig * dps_rcp(v_abs);
            dv_re += t * v_re;
            dv_im += t * v_im;
        }
    } else {
        gp[DPS_S_EFD0] += d_efd;
    }
    if (CTRL && has_pss) {
        const double xw = x[DPS_X_PSS_W], x1 = x[DPS_X_PSS_LL1], x2 = x[DPS_X_PSS_LL2];
        if (fed) {
            const double hl = par(DPS_S_S_HLIM);
            const double d_o2 = d_vpss * clamp_dx(o2, -hl, hl);
            gp[DPS_S_S_HLIM] += d_vpss * (clamp_dhi(o2, -hl, hl) - clamp_dlo(o2, -hl, hl));
            double d_o1 = 0.0, d_w = 0.0;
            {   // f12 = inv_t4*(o1 - x2)
                const double a = df[DPS_X_PSS_LL2];
                gp[DPS_S_S_INV_T4] += a * (o1 - x2);
                const double t = a * par(DPS_S_S_INV_T4);
                d_o1 += t;
                dx[DPS_X_PSS_LL2] -= t;
            }
            gp[DPS_S_S_R2] += d_o2 * o1;                                // o2 = r2*o1 + omr2*x2
            d_o1 += d_o2 * par(DPS_S_S_R2);
            gp[DPS_S_S_OMR2] += d_o2 * x2;
            dx[DPS_X_PSS_LL2] += d_o2 * par(DPS_S_S_OMR2);
            {   // f11 = inv_t3*(w - x1)
                const double a = df[DPS_X_PSS_LL1];
                gp[DPS_S_S_INV_T3] += a * (w - x1);
                const double t = a * par(DPS_S_S_INV_T3);
                d_w += t;
                dx[DPS_X_PSS_LL1] -= t;
            }
            gp[DPS_S_S_R1] += d_o1 * w;                                 // o1 = r1*w + omr1*x1
            d_w += d_o1 * par(DPS_S_S_R1);
            gp[DPS_S_S_OMR1] += d_o1 * x1;
            dx[DPS_X_PSS_LL1] += d_o1 * par(DPS_S_S_OMR1);
            d_w += df[DPS_X_PSS_W];                                     // f10 = w
            gp[DPS_S_S_INV_TW] += d_w * (par(DPS_S_S_KW) * omega - xw); // w = inv_tw*(kw*omega - xw)
            const double t = d_w * par(DPS_S_S_INV_TW);
            gp[DPS_S_S_KW] += t * omega;
            d_omega += t * par(DPS_S_S_KW);
            dx[DPS_X_PSS_W] -= t;
        } else {
            gp[DPS_S_S_INV_TW] += df[DPS_X_PSS_W] * (-xw);
            dx[DPS_X_PSS_W] -= df[DPS_X_PSS_W] * par(DPS_S_S_INV_TW);
            gp[DPS_S_S_INV_T3] += df[DPS_X_PSS_LL1] * (-x1);
            dx[DPS_X_PSS_LL1] -= df[DPS_X_PSS_LL1] * par(DPS_S_S_INV_T3);
            gp[DPS_S_S_INV_T4] += df[DPS_X_PSS_LL2] * (-x2);
            dx[DPS_X_PSS_LL2] -= df[DPS_X_PSS_LL2] * par(DPS_S_S_INV_T4);
        }
    }
    dx[DPS_X_OMEGA] += d_omega;

    // ---- algebraic chain, synchronous_machine.py:273-277 -----------------------------------------------------------
    double d_ire = d_id * sd + d_iq * cd;                 // i_d = i_re s - i_im c ; i_q = i_re c + i_im s
    double d_iim = -d_id * cd + d_iq * sd;
    dsd += d_id * i_re + d_iq * i_im;
    dcd += -d_id * i_im + d_iq * i_re;
    dv_re += d_pe * i_re;                                 // p_e = v_re i_re + v_im i_im
    dv_im += d_pe * i_im;
    d_ire += d_pe * v_re;
    d_iim += d_pe * v_im;
    gp[DPS_S_INV_XDST] += d_ire * (e_im - v_im) - d_iim * (e_re - v_re);
    const double d_eim = d_ire * ixd;                     // i_re = (e_im - v_im) ixd ; i_im = -(e_re - v_re) ixd
    dv_im -= d_ire * ixd;
    const double d_ere = -d_iim * ixd;
    dv_re += d_iim * ixd;
    dx[DPS_X_EQST] += d_ere * cd + d_eim * sd;            // e_re = eq c + ed s ; e_im = eq s - ed c
    dx[DPS_X_EDST] += d_ere * sd - d_eim * cd;
    dcd += d_ere * eqst - d_eim * edst;
    dsd += d_ere * edst + d_eim * eqst;
}

// Transpose of injection(): given the cotangent (di_re, di_im) of this generator's Norton current.
template <class PR, class GP>
__device__ __forceinline__ void injection_vjp(const double* x, const PR& par, double sd, double cd, double di_re,
                                              double di_im, double* dx, GP& gp, double& dsd, double& dcd) {
    const double ixd = par(DPS_S_INV_XDST), ixq = par(DPS_S_INV_XQST), snr = par(DPS_S_SNR);
    const double a = x[DPS_X_EQST] * ixd, b = x[DPS_X_EDST] * ixq;
    gp[DPS_S_SNR] += di_re * (a * sd - b * cd) + di_im * (-(a * cd) - b * sd);
    const double t_re = di_re * snr, t_im = di_im * snr;
    const double da = t_re * sd - t_im * cd;
    const double db = -t_re * cd - t_im * sd;
    dsd += t_re * a - t_im * b;
    dcd += -t_re * b - t_im * a;
    dx[DPS_X_EQST] += da * ixd;
    gp[DPS_S_INV_XDST] += da * x[DPS_X_EQST];
    dx[DPS_X_EDST] += db * ixq;
    gp[DPS_S_INV_XQST] += db * x[DPS_X_EDST];
}

}  // namespace dps
