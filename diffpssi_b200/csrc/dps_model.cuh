// dps_model.cuh -- the per-generator stage evaluation shared by the forward stepper and the adjoint.
//
// Real-fp64 restatement of what one stage of solver.step() does for one generator of one trajectory
// (reference file:line in the comments below).  The reference carries complex128 tensors whose imaginary
// part is identically zero for every *state* (SURVEY.md fact 3); only bus voltages / currents / Z are complex,
// and those are handled here as explicit (re, im) pairs.
#pragma once
#include <math.h>
#include "../../include/dps.h"

namespace dps {

// ---- math primitives ----------------------------------------------------------------------------------------------
// Device versions are branch-free (no slow-path subroutine, no local-memory scratch): the kernel is bound by the FP64
// pipe and by issue slots, and the CUDA library versions of sincos / division / sqrt each carry a divergent slow path.
// Accuracy is ~1 ulp, far inside the 1e-9 p.u. parity budget.  The host versions (used only by the CPU emulation in
// tests/host_emul) call libm.
#ifdef __CUDACC__
// fdlibm kernel coefficients (k_sin.c / k_cos.c) and the Cody-Waite split of pi/2 (e_rem_pio2.c); kept in constant
// memory so each DFMA takes its coefficient as a c[bank][offset] operand instead of two UMOVs.
static __constant__ double dps_trig_c[16] = {
    6.36619772367581382433e-01, 6755399441055744.0, 1.57079632673412561417e+00, 6.07710050650619224932e-11,
    1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06, -1.98412698298579493134e-04,
    8.33333333332248946124e-03, -1.66666666666666324348e-01,
    -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07, 2.48015872894767294178e-05,
    -1.38888888888741095749e-03, 4.16666666666666019037e-02};
#endif
#ifdef __CUDA_ARCH__
// rcp.approx / rsqrt.approx deliver >= 20 good bits; one Newton step (>= 40 bits) followed by one residual
// correction of the final quotient / root (which doubles the accuracy again) is correctly rounded in all but a
// vanishing fraction of cases (checked against libm by tests/test_cuda_parity.py::test_device_math).
__device__ __forceinline__ double dps_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    const double e2 = fma(-x, r, 1.0);
    return fma(r, e2, r);
}
__device__ __forceinline__ double dps_div(double a, double b) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    const double e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    const double q = a * r;
    return fma(fma(-q, b, a), r, q);             // residual correction
}
__device__ __forceinline__ double dps_sqrt(double s) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
    const double h = 0.5 * y;
    const double e = fma(-s * y, h, 0.5);          // 0.5 - 0.5*s*y^2
    y = fma(y, e, y);
    const double r = s * y;
    const double res = fma(fma(-r, r, s), 0.5 * y, r);
    // s is a sum of squares: an integer test of its high word (normal and positive) replaces the FP64 compare
    return (__double2hiint(s) >= 0x00100000) ? res : 0.0;
}
// a / 6 (RK4 combination, solvers.py:214): product with the rounded reciprocal plus one residual correction -- correctly
// rounded except for a vanishing fraction of inputs (then off by one ulp); the IEEE division sequence is ~25 instructions.
__device__ __forceinline__ double dps_div6(double a) {
    const double r = 1.0 / 6.0;
    const double q = a * r;
    return fma(fma(-q, 6.0, a), r, q);
}
// sqrt of a strictly positive normal number: no zero guard (the caller adds a 1e-300 floor for free inside an FMA)
__device__ __forceinline__ double dps_sqrt_pos(double s) {
    // coupled (Goldschmidt) iteration: g -> sqrt(s), h -> 1/(2 sqrt(s)); five dependent FP64 levels after the MUFU
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
    const double g = s * y, h = 0.5 * y;
    const double e = fma(-g, h, 0.5);
    const double g1 = fma(g, e, g), h1 = fma(h, e, h);
    return fma(fma(-g1, g1, s), h1, g1);         // residual correction
}
// sin/cos with 2-term Cody-Waite reduction by pi/2 and the fdlibm kernel polynomials on [-pi/4, pi/4].
// Rotor angles stay far below the 1e6 rad where the 2-term reduction starts to lose bits.
__device__ __forceinline__ void dps_sincos(double x, double* sp, double* cp) {
    const double* C = dps_trig_c;
    const double t = fma(x, C[0], C[1]);
    const int q = __double2loint(t);
    const double k = t - C[1];
    double r = fma(-k, C[2], x);
    r = fma(-k, C[3], r);
    // Estrin evaluation of the two fdlibm kernel polynomials: three dependent levels instead of Horner's five (the
    // kernel is bound by FP64 dependency latency, not by the pipe's throughput -- profiles/README.md)
    const double z = r * r;
    const double z2 = z * z;
    const double sa = fma(z, C[4], C[5]), sb = fma(z, C[6], C[7]), sc = fma(z, C[8], C[9]);
    const double ca = fma(z, C[10], C[11]), cb = fma(z, C[12], C[13]), cc = fma(z, C[14], C[15]);
    const double ps = fma(fma(sa, z2, sb), z2, sc);
    const double pc = fma(fma(ca, z2, cb), z2, cc);
    const double sr = fma(z * r, ps, r);
    const double cr = fma(z2, pc, fma(z, -0.5, 1.0));
    const double s = (q & 1) ? cr : sr;
    const double c = (q & 1) ? sr : cr;
    // quadrant signs are XORed into the high words (integer pipe) instead of FP64 negations
    *sp = __hiloint2double(__double2hiint(s) ^ ((q & 2) << 30), __double2loint(s));
    *cp = __hiloint2double(__double2hiint(c) ^ (((q + 1) & 2) << 30), __double2loint(c));
}
#else
inline double dps_rcp(double x) { return 1.0 / x; }
inline double dps_div(double a, double b) { return a / b; }
inline double dps_sqrt(double s) { return sqrt(s); }
inline double dps_sqrt_pos(double s) { return sqrt(s); }
inline double dps_div6(double a) { return a / 6.0; }
inline void dps_sincos(double x, double* s, double* c) { sincos(x, s, c); }
#endif

// Recorder store.  A warp of the G <= 16 stepper holds 32/G trajectories, so it writes 8*(32/G)-byte pieces of each
// recorder row and the neighbouring warps complete the 32-byte DRAM sectors a few steps earlier or later.  A plain
// (write-back) store lets L2 merge the pieces; the streaming hint st.global.cs evicted them half-filled and cost a DRAM
// read-modify-write per piece (14.1 GB of traffic per launch instead of the algorithmic 10.5 GB, profiles/README.md).
#ifdef __CUDACC__
__device__ __forceinline__ void dps_rec_store(double* p, double v) {
#ifdef DPS_REC_STREAMING
    __stcs(p, v);
#else
    *p = v;
#endif
}
#endif

// Reads constant `slot` of the generator this thread works on.  UP = uniform table in shared memory,
// VARY=true adds the per-sim override path.
template <bool VARY>
struct ParamReader {
    const double* up;        // shared memory, [DPS_N_SLOTS] of this generator
    const short*  sv;        // shared memory, [DPS_N_SLOTS] row in vpar or -1
    const double* vpar;      // global [n_vary][P]
    long long     P, s;
    __device__ __forceinline__ double operator()(int slot) const {
        if (VARY) {
            const int vi = sv[slot];
            if (vi >= 0) return __ldg(vpar + (long long)vi * P + s);
        }
        return up[slot];
    }
};

// Fast-path reader: the DPS_N_REG_SLOTS leading constants of this (trajectory, generator) -- the 15 machine constants
// (what parameter identification varies per sim: H, D, X_d, X_q, X_d', X_q', ...) and the three controller set points
// initialize() derives from them -- live in registers; the remaining controller constants are uniform and come
// from shared memory as warp-broadcast loads at compile-time offsets.  `slot` is a literal at every call site, so the
// branch folds away.
struct ParamRegs {
    const double* pm;        // registers, [DPS_N_REG_SLOTS]
    const double* up;        // shared memory, [DPS_N_SLOTS] uniform constants of this generator
    __device__ __forceinline__ double operator()(int slot) const {
        return (slot < DPS_N_REG_SLOTS) ? pm[slot] : up[slot];
    }
};

// Fastest reader (PMODE 3): as ParamRegs, but the controller constants are identical for every generator and every
// trajectory, live in the kernel's parameter block (constant bank) and fold into the FP64 instructions as c[0][off]
// operands -- no load instruction, no shared-memory wavefront.
struct ParamConst {
    const double* pm;        // registers, [DPS_N_REG_SLOTS]
    const double* cp;        // kernel parameter block, [DPS_N_SLOTS]
    __device__ __forceinline__ double operator()(int slot) const {
        return (slot < DPS_N_REG_SLOTS) ? pm[slot] : cp[slot];
    }
};

struct Algebraics {            // values the reference keeps on the machine object after a stage
    double p_e, i_d, i_q, v_re, v_im, e_fd, p_m;
};

// Norton current injection, synchronous_machine.py:256-260:
//   i_d = e_q''/(j x_d'') e^{j delta},  i_q = e_d''/(j x_q'') e^{j(delta - pi/2)},  I = (i_d + i_q) S_n/S_sys
// with e^{j delta} = c + j s and e^{j(delta-pi/2)} = s - j c this is
//   I = snr * [ (a s - b c) - j (a c + b s) ],  a = e_q''/x_d'',  b = e_d''/x_q''
template <class PR>
__device__ __forceinline__ void injection(const double* x, const PR& par, double sd, double cd,
                                          double& i_re, double& i_im) {
    // the base change S_n/S_sys is folded into a and b (two multiplications fewer than scaling both components)
    const double snr = par(DPS_S_SNR);
    const double a = x[DPS_X_EQST] * (par(DPS_S_INV_XDST) * snr);
    const double b = x[DPS_X_EDST] * (par(DPS_S_INV_XQST) * snr);
    i_re = a * sd - b * cd;
    i_im = -(a * cd + b * sd);
}

__device__ __forceinline__ double clampd(double v, double lo, double hi) {
    // blocks.py:181, 241: minimum(maximum(x.real, lo), hi).  Written as compare+select: the library fmin/fmax expand to
    // ~8 integer instructions each for their NaN rules (a NaN state stays NaN here, as it does in torch).
    const double m = (v < lo) ? lo : v;
    return (m > hi) ? hi : m;
}

// PT1Limited (blocks.py:138-150, 170-182): output = clamp(state, lo, hi), and the anti-windup gate freezes the
// derivative when (state <= lo and dx < 0) or (state >= hi and dx > 0).  The two FP64 comparisons state<=lo / state>=hi
// serve both (for lo <= hi, `state<=lo ? lo : state>=hi ? hi : state` equals minimum(maximum(state, lo), hi)); the sign
// tests of dx are integer tests on its high word (a -0.0 derivative is frozen to +0.0, which is the same number).
struct Pt1Lim {
    bool le, ge;
    double out;
    __device__ __forceinline__ Pt1Lim(double state, double lo, double hi) {
        le = state <= lo;
        ge = state >= hi;
        out = le ? lo : (ge ? hi : state);
    }
    __device__ __forceinline__ double gate(double dx) const {
#ifdef __CUDA_ARCH__
        // one predicate, one masking of the two words (the short-circuit form compiled into a cascade of six selects)
        const int hw = __double2hiint(dx), lw = __double2loint(dx);
        const bool neg = hw < 0;
        const bool pos = (hw >= 0) & ((hw | lw) != 0);
        const int keep = ((le & neg) | (ge & pos)) ? 0 : -1;
        return __hiloint2double(hw & keep, lw & keep);
#else
        const bool neg = dx < 0.0, pos = dx > 0.0;
        return ((le && neg) || (ge && pos)) ? 0.0 : dx;
#endif
    }
};

// Anti-windup gate of PT1Limited.differential, blocks.py:145-149: `<=`/`>=` on the state, strict on the derivative
__device__ __forceinline__ double windup_gate(double state, double dx, double lo, double hi) {
    const bool stop = ((state <= lo) && (dx < 0.0)) || ((state >= hi) && (dx > 0.0));
    return stop ? 0.0 : dx;
}

// One stage for one generator, given its bus voltage (v_re, v_im): update_internal_vars
// (synchronous_machine.py:271-277) + differential() (synchronous_machine.py:148-172) with the controller
// get_output()/differential() pairs it calls.  Writes f[0..12] (zero for absent controllers).
//
// LIM = true applies the PT1Limited clamp / anti-windup selects (blocks.py:145-149, 181).  LIM = false is the same code
// without them, valid when no PT1Limited block sits at a limit (it then passes its state and derivative through unchanged):
// stage_derivatives<CTRL, VOTE = true> tests the limits of all 32 lanes with one warp vote (they are functions of the
// states only) and takes the select-free body when no lane needs them -- the common case outside fault transients.
// Each body is one straight-line block, so the scheduler can interleave its independent chains freely.
template <bool CTRL, bool LIM, class PR>
__device__ __forceinline__ void stage_body(const double* x, const PR& par, int ctrl, double fn,
                                           double sd, double cd, double v_re, double v_im,
                                           double* f, Algebraics& alg) {
    const double omega = x[DPS_X_OMEGA];
    const double eqst = x[DPS_X_EQST], edst = x[DPS_X_EDST];
    // e'' = e_q'' e^{j delta} + e_d'' e^{j(delta - pi/2)}                              :273
    // (every multiply-add below is an explicit fma: the LIM / non-LIM instantiations then round identically, so a
    // trajectory does not depend on which other trajectories share its warp)
    const double e_re = fma(eqst, cd, edst * sd);
    const double e_im = fma(eqst, sd, -(edst * cd));
    // i_gen = (e'' - v)/(j x_d'')   (x_d'' on both axes, as the reference)             :275
    const double inv_xdst = par(DPS_S_INV_XDST);
    const double i_re = (e_im - v_im) * inv_xdst;
    const double i_im = -((e_re - v_re) * inv_xdst);
    const double p_e = fma(v_re, i_re, v_im * i_im);                                 // :274
    // (i_d + j i_q) = i_gen e^{j(pi/2 - delta)} = i_gen (s + j c)                      :276-277
    const double i_d = fma(i_re, sd, -(i_im * cd));
    const double i_q = fma(i_re, cd, i_im * sd);

    double e_fd = par(DPS_S_EFD0);      // held constant without an exciter (synchronous_machine.py:314)
    double p_m = par(DPS_S_PM0);        // held constant without a governor (synchronous_machine.py:313)
#pragma unroll
    for (int k = 6; k < DPS_N_STATES; ++k) f[k] = 0.0;

    if (CTRL) {
        constexpr bool limiting = LIM;
        double o2 = 0.0, hl = 0.0;
        if (ctrl & DPS_CTRL_PSS) {
            // STAB1.get_output is only called when an exciter exists (synchronous_machine.py:148-150); without
            // one the block inputs keep their initial 0 and only the decay terms remain.
            const bool fed = (ctrl & DPS_CTRL_AVR) != 0;
            const double xw = x[DPS_X_PSS_W], x1 = x[DPS_X_PSS_LL1], x2 = x[DPS_X_PSS_LL2];
            // Washout: output == derivative, blocks.py:382, 413
            const double w = par(DPS_S_S_INV_TW) * fma(par(DPS_S_S_KW), (fed ? omega : 0.0), -xw);
            const double in1 = fed ? w : 0.0;
            const double o1 = fma(par(DPS_S_S_R1), in1, par(DPS_S_S_OMR1) * x1);     // blocks.py:319
            const double in2 = fed ? o1 : 0.0;
            o2 = fma(par(DPS_S_S_R2), in2, par(DPS_S_S_OMR2) * x2);
            f[DPS_X_PSS_W] = w;
            f[DPS_X_PSS_LL1] = par(DPS_S_S_INV_T3) * (in1 - x1);                     // blocks.py:290
            f[DPS_X_PSS_LL2] = par(DPS_S_S_INV_T4) * (in2 - x2);
            hl = par(DPS_S_S_HLIM);
        }
        double xe = 0.0, emin = 0.0, emax = 0.0, xg = 0.0, vmin = 0.0, vmax = 0.0;
        if (ctrl & DPS_CTRL_AVR) {
            xe = x[DPS_X_AVR_PT1]; emin = par(DPS_S_A_EMIN); emax = par(DPS_S_A_EMAX);
        }
        if (ctrl & DPS_CTRL_GOV) {
            xg = x[DPS_X_GOV_PT1]; vmin = par(DPS_S_G_VMIN); vmax = par(DPS_S_G_VMAX);
        }
        double v_pss = 0.0;
        // the stabiliser's output limiter is applied unconditionally (two selects): with the stock H_lim = 0.03 it
        // saturates during most of a transient, and routing it through the vote would send 40 % of the warps down the
        // slow body for the sake of a clamp
        if (ctrl & DPS_CTRL_PSS) v_pss = clampd(o2, -hl, hl);                        // blocks.py:241
        if (ctrl & DPS_CTRL_AVR) {
            // SEXS.get_output(|v_bb| - v_pss), exciters.py:111-114
            // |v|: the 1e-300 floor (far below one ulp of any |v|^2 > 1e-284) replaces the zero guard of the root
            const double v_abs = dps_sqrt_pos(fma(v_re, v_re, fma(v_im, v_im, 1e-300)));
            // u = v_set - (|v| - v_pss) + bias (exciters.py:113), associated so that only ONE addition follows the root:
            // the set-point terms do not depend on the network solve and are summed while it is in flight
            const double c0 = (ctrl & DPS_CTRL_PSS) ? ((par(DPS_S_A_VSET) + par(DPS_S_A_BIAS)) + v_pss)
                                                    : (par(DPS_S_A_VSET) + par(DPS_S_A_BIAS));
            const double xll = x[DPS_X_AVR_LL];
            const double oll = par(DPS_S_A_OMR) * xll;
            const double u = c0 - v_abs;
            const double o = fma(par(DPS_S_A_R), u, oll);
            f[DPS_X_AVR_LL] = par(DPS_S_A_INV_TB) * (u - xll);
            const double dxe = par(DPS_S_A_INV_TE) * fma(par(DPS_S_A_K), o, -xe);    // blocks.py:143
            e_fd = xe;
            f[DPS_X_AVR_PT1] = dxe;
            if (limiting) {
                const Pt1Lim lim(xe, emin, emax);
                e_fd = lim.out;
                f[DPS_X_AVR_PT1] = lim.gate(dxe);
            }
        }
        if (ctrl & DPS_CTRL_GOV) {
            // TGOV1.get_output(omega), governors.py:115-123
            const double u = par(DPS_S_G_PREF) - omega;
            const double xll = x[DPS_X_GOV_LL];
            const double dxg = par(DPS_S_G_INV_T1) * fma(par(DPS_S_G_GAIN), u, -xg);
            double gs = xg;
            f[DPS_X_GOV_PT1] = dxg;
            if (limiting) {
                const Pt1Lim lim(xg, vmin, vmax);
                gs = lim.out;
                f[DPS_X_GOV_PT1] = lim.gate(dxg);
            }
            const double o = fma(par(DPS_S_G_R), gs, par(DPS_S_G_OMR) * xll);
            p_m = fma(-par(DPS_S_G_DT), omega, o);
            f[DPS_X_GOV_LL] = par(DPS_S_G_INV_T3) * (gs - xll);
        }
    }

    // synchronous_machine.py:158-164
    const double t_m = dps_div(p_m, 1.0 + omega);
    f[DPS_X_OMEGA] = par(DPS_S_INV_2H) * (fma(-par(DPS_S_D), omega, t_m) - p_e);    // (t_m - p_e - D omega)/(2H)
    f[DPS_X_DELTA] = omega * ((2.0 * 3.141592653589793) * fn);
    f[DPS_X_EQT] = par(DPS_S_INV_TD0T) * fma(-i_d, par(DPS_S_XD_XDT), e_fd - x[DPS_X_EQT]);
    f[DPS_X_EDT] = par(DPS_S_INV_TQ0T) * fma(i_q, par(DPS_S_XQ_XQT), -x[DPS_X_EDT]);
    f[DPS_X_EQST] = par(DPS_S_INV_TD0ST) * fma(-i_d, par(DPS_S_XDT_XDST), x[DPS_X_EQT] - eqst);
    f[DPS_X_EDST] = par(DPS_S_INV_TQ0ST) * fma(i_q, par(DPS_S_XQT_XQST), x[DPS_X_EDT] - edst);

    alg.p_e = p_e; alg.i_d = i_d; alg.i_q = i_q; alg.v_re = v_re; alg.v_im = v_im; alg.e_fd = e_fd; alg.p_m = p_m;
}

// True when a PT1Limited block of this generator sits at / beyond one of its limits at state x: the only situation in
// which the LIM selects of stage_body change a value (the stabiliser's output clamp is always applied).
template <class PR>
__device__ __forceinline__ bool stage_at_limit(const double* x, const PR& par, int ctrl) {
    bool at = false;
    if (ctrl & DPS_CTRL_AVR) {
        const double xe = x[DPS_X_AVR_PT1];
        at = (xe <= par(DPS_S_A_EMIN)) || (xe >= par(DPS_S_A_EMAX));
    }
    if (ctrl & DPS_CTRL_GOV) {
        const double xg = x[DPS_X_GOV_PT1];
        at = at || (xg <= par(DPS_S_G_VMIN)) || (xg >= par(DPS_S_G_VMAX));
    }
    return at;
}

// VOTE (device only): the caller guarantees that all 32 lanes of the warp are in the call.
template <bool CTRL, bool VOTE = false, class PR>
__device__ __forceinline__ void stage_derivatives(const double* x, const PR& par, int ctrl, double fn,
                                                  double sd, double cd, double v_re, double v_im,
                                                  double* f, Algebraics& alg) {
#ifdef __CUDA_ARCH__
    if (CTRL && VOTE) {
        if (__any_sync(0xffffffffu, stage_at_limit(x, par, ctrl)))
            stage_body<CTRL, true>(x, par, ctrl, fn, sd, cd, v_re, v_im, f, alg);
        else
            stage_body<CTRL, false>(x, par, ctrl, fn, sd, cd, v_re, v_im, f, alg);
        return;
    }
#endif
    stage_body<CTRL, true>(x, par, ctrl, fn, sd, cd, v_re, v_im, f, alg);
}

}  // namespace dps
