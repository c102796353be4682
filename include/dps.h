/*
 * dps.h -- C ABI of the B200-native DiffPSSi stepping engine (libdps_b200.so).
 *
 * The reference (georgkordowich/diffpssi 0.0.3) has no FFI: its hot path is the Python loop
 * ``for t in self.time`` of PowerSystemSimulation.run() (src/diffpssi/power_sim_lib/simulator.py:523-543)
 * which calls solver.step() (solvers.py:26/75/143) and the model methods.  This header draws the boundary
 * SURVEY.md section 8(b) proposes: *after* initialize() (simulator.py:505) and *around the whole loop*
 * (simulator.py:523-546).  One dps_forward() call integrates every step of every trajectory and fills the
 * recorder; what each entry point replaces is cited at its declaration.
 *
 * Conventions
 *   - plain C: pointers, sizes, POD structs.  No C++/torch types cross the boundary.
 *   - every array argument is a DEVICE pointer unless it is a member of dps_topology (host memory, copied
 *     by dps_plan_create) or documented as host.
 *   - arrays are fp64, structure-of-arrays with the batch (parallel_sims) index innermost:
 *         a[i][p]  ==  a[i * P + p]
 *   - the caller owns every buffer; the library allocates nothing on the device.
 *   - all launches are asynchronous on `stream` (a cudaStream_t passed as void*); no hidden synchronisation.
 *   - return value 0 = success; otherwise dps_last_error() holds a message (thread local).
 *   - NaN/Inf in the data is not an error: diverged trajectories propagate like they do in the reference.
 */
#ifndef DPS_H
#define DPS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPS_ABI_VERSION 4

/* ---- per-generator state layout (canonical: absent controllers keep zero slots) ----------------------------
 * reference order: machine(6) | exciter(lead-lag, PT1) | governor(PT1, lead-lag) | PSS(washout, LL1, LL2)
 * (synchronous_machine.py:219-228, exciters.py:99-100, governors.py:102-103, stabilizers.py:102-104)        */
#define DPS_N_STATES 13
enum dps_state {
    DPS_X_OMEGA = 0, DPS_X_DELTA, DPS_X_EQT, DPS_X_EDT, DPS_X_EQST, DPS_X_EDST,
    DPS_X_AVR_LL, DPS_X_AVR_PT1, DPS_X_GOV_PT1, DPS_X_GOV_LL, DPS_X_PSS_W, DPS_X_PSS_LL1, DPS_X_PSS_LL2
};

/* ---- recordable per-generator fields: the 13 states (post-step values) followed by the algebraics the
 * reference keeps on the machine object, which hold the values of the LAST stage evaluation of the step
 * (simulator.py:540-543; synchronous_machine.py:271-277, 148-156)                                           */
enum dps_field {
    DPS_F_PE = 13, DPS_F_ID, DPS_F_IQ, DPS_F_VRE, DPS_F_VIM, DPS_F_EFD, DPS_F_PM,
    DPS_N_FIELDS
};

/* ---- per-generator constants ("slots").  Derived on the host from the model parameters with the same
 * expressions the reference evaluates inside the loop (e.g. inv_2h = 1/(2*h), synchronous_machine.py:159), so
 * multiplying by them is bit-identical to the reference's ``1 / T * (...)``.                                 */
#define DPS_N_SLOTS 42
#define DPS_N_REG_SLOTS 18   /* leading slots the fast kernel keeps in registers (may differ per trajectory) */
enum dps_slot {
    /* machine: synchronous_machine.py:158-164, 256-277 */
    DPS_S_INV_2H = 0, DPS_S_D, DPS_S_INV_TD0T, DPS_S_XD_XDT, DPS_S_INV_TQ0T, DPS_S_XQ_XQT,
    DPS_S_INV_TD0ST, DPS_S_XDT_XDST, DPS_S_INV_TQ0ST, DPS_S_XQT_XQST,
    DPS_S_INV_XDST, DPS_S_INV_XQST, DPS_S_SNR, DPS_S_PM0, DPS_S_EFD0,
    /* controller constants that initialize() derives from the machine state (exciters.py:137-138, simulator.py:319,
       governors.py:145): they differ per trajectory whenever a machine parameter does */
    DPS_S_A_VSET, DPS_S_A_BIAS, DPS_S_G_PREF,
    /* SEXS: exciters.py:102-115, blocks.py:143-150 (PT1Limited), 290/319 (LeadLag) */
    DPS_S_A_R, DPS_S_A_OMR, DPS_S_A_INV_TB, DPS_S_A_INV_TE, DPS_S_A_K, DPS_S_A_EMIN, DPS_S_A_EMAX,
    /* TGOV1: governors.py:105-123 */
    DPS_S_G_INV_T1, DPS_S_G_GAIN, DPS_S_G_VMIN, DPS_S_G_VMAX, DPS_S_G_R, DPS_S_G_OMR, DPS_S_G_INV_T3,
    DPS_S_G_DT,
    /* STAB1: stabilizers.py:106-121, blocks.py:382/413 (Washout), 241 (Limiter) */
    DPS_S_S_INV_TW, DPS_S_S_KW, DPS_S_S_R1, DPS_S_S_OMR1, DPS_S_S_INV_T3, DPS_S_S_R2, DPS_S_S_OMR2,
    DPS_S_S_INV_T4, DPS_S_S_HLIM
};

enum dps_solver { DPS_EULER = 0, DPS_HEUN = 1, DPS_RK4 = 2 };      /* solvers.py:226-230 */
enum dps_ctrl   { DPS_CTRL_AVR = 1, DPS_CTRL_GOV = 2, DPS_CTRL_PSS = 4 };

enum dps_channel_kind { DPS_CH_GEN = 0, DPS_CH_BUS_VRE = 1, DPS_CH_BUS_VIM = 2 };

typedef struct {
    int32_t kind;     /* dps_channel_kind */
    int32_t index;    /* generator index (DPS_CH_GEN) or observed-row index (DPS_CH_BUS_*)  */
    int32_t field;    /* dps_state / dps_field for DPS_CH_GEN, ignored otherwise             */
} dps_channel;

/* Host-side description of one run; copied by dps_plan_create (all pointers are HOST memory). */
typedef struct {
    int32_t abi_version;      /* DPS_ABI_VERSION */
    int32_t n_gen;            /* generators, in bus order then model order (solvers.py:84-86) */
    int32_t solver;           /* dps_solver */
    int32_t n_steps;          /* T = len(np.arange(0, sim_time, dt)) (simulator.py:78) */
    double  dt;
    double  fn;               /* system frequency; d(delta)/dt = omega*2*pi*fn (synchronous_machine.py:160) */
    const int32_t* ctrl;      /* [n_gen] dps_ctrl bit flags */
    int32_t n_obs;            /* rows of the reduced impedance matrix: the n_gen generator buses first (generator
                                 order), then any extra buses whose voltage is recorded */
    int32_t z_rows;           /* row count  of each Zsub configuration as stored (>= n_obs)  */
    int32_t z_cols;           /* column count of each Zsub configuration as stored (>= n_gen) */
    int32_t n_cfg;            /* network configurations; 0 = no fault.  One per distinct set of faulted buses. */
    const uint8_t* step_cfg;  /* [n_steps] configuration active during step k (ScEvent timing, static_models.py:46,
                                 simulator.py:524-531, evaluated by the host on the reference's own time grid) */
    int32_t z_per_sim;        /* 0: Zsub shared by the batch; 1: one Zsub per sim */
    const int32_t* slot_vary; /* [n_gen*DPS_N_SLOTS]: -1 = uniform (value in upar), else row of vpar */
    int32_t n_vary;           /* rows of vpar */
    int32_t n_rec;
    const dps_channel* rec;   /* [n_rec] recorder selectors */
    int32_t ckpt_stride;      /* 0 = none; else states are checkpointed every ckpt_stride steps (adjoint) */
    const double* common_ctrl;/* optional HOST [DPS_N_SLOTS]: when every generator carries all three controllers with
                                 batch-uniform constants that are also identical between generators, their values
                                 (slots >= DPS_N_REG_SLOTS are read); the stepper then takes them as immediate
                                 constant-bank operands instead of shared-memory loads.  NULL otherwise. */
} dps_topology;

typedef struct dps_plan dps_plan;

/* Build / destroy a plan (host object; validates the topology, picks the kernel specialisation). */
int dps_plan_create(const dps_topology* topo, dps_plan** out);
int dps_plan_destroy(dps_plan* plan);

/* Number of checkpoints a forward call writes for this plan (0 if ckpt_stride == 0). */
int64_t dps_n_checkpoints(const dps_plan* plan);

/* Bytes of scratch dps_backward needs for P trajectories. */
size_t dps_workspace_bytes(const dps_plan* plan, int64_t P);

/* Bytes of the OPTIONAL scratch of dps_forward / dps_forward_range for P trajectories (a few bytes per trajectory;
 * 0 when the plan's stepper has no use for it).  With it -- and with x_final / x_out non-NULL -- a launch whose blocks
 * would make only a few "waves" over the SMs runs as a persistent grid that pulls (block of trajectories, chunk of
 * steps) work items from a queue, so that a partly filled last wave does not idle most of the GPU for the whole run
 * (e.g. the 8192-trajectory shards of the 8-GPU configuration).  Results are bit-identical with and without. */
size_t dps_forward_workspace_bytes(const dps_plan* plan, int64_t P);

/*
 * Fault re-reduction on the device.  Replaces the per-step
 *     Y = inv(Z); Y[:, b, b] = 1e6; Z = inv(Y)            (simulator.py:526-528)
 * by one rank-1 (Sherman-Morrison) update per faulted bus, computed once per configuration:
 *     Z' = Z - f * Z[:,b] Z[b,:],   f = c / (1 + c Z_bb),   c = y_fault - Y_bb
 * (the diagonal is *set*, not incremented), and gathers the rows/columns the stepper needs.
 *
 *   Bz            1 (network shared by the batch) or P
 *   z_re, z_im    [n_bus][n_bus][Bz]   full Z = inv(Y_dyn) of the un-faulted network
 *   yd_re, yd_im  [n_bus][Bz]          diagonal of Y_dyn
 *   obs_bus       HOST [n_obs]  bus index of each observed row (generator buses first)
 *   gen_bus       HOST [n_gen]  bus index of each generator column
 *   cfg_ptr       HOST [n_cfg+1], cfg_bus HOST [cfg_ptr[n_cfg]]: faulted buses of each configuration (CSR)
 *   zsub          out [n_cfg][z_rows][z_cols][2][Bz] with z_rows = n_obs + m_max, z_cols = n_gen + m_max,
 *                 m_max = max faulted buses in one configuration; entries beyond [n_obs][n_gen] are scratch.
 */
int dps_network_reduce(int64_t Bz, int32_t n_bus,
                       const double* z_re, const double* z_im, const double* yd_re, const double* yd_im,
                       int32_t n_obs, const int32_t* obs_bus, int32_t n_gen, const int32_t* gen_bus,
                       int32_t n_cfg, const int32_t* cfg_ptr, const int32_t* cfg_bus,
                       double y_fault, double* zsub, void* stream);

/*
 * The time loop.  Replaces simulator.py:523-543 for all P trajectories:
 * fault switching (ScEvent, via step_cfg), solver.step() (solvers.py) with the model equations
 * (synchronous_machine.py:138-277, blocks.py, exciters.py, governors.py, stabilizers.py), the network solve
 * V = Z I (solvers.py:83/102) restricted to generator columns / observed rows, and the recorder
 * (simulator.py:540-546).
 *
 *   x0       [n_gen][DPS_N_STATES][P]       initial states (canonical layout)
 *   upar     [n_gen][DPS_N_SLOTS]           uniform constants
 *   vpar     [n_vary][P]                    per-sim constants (may be NULL if n_vary == 0)
 *   zsub     [n_cfg][z_rows][z_cols][2][Bz] reduced impedance, Bz = z_per_sim ? P : 1
 *   rec      out [n_rec][n_steps][P]        == the reference's [batch, timestep, value] result viewed as
 *                                           rec.permute(2,1,0) (same memory order as simulator.py:546)
 *   x_final  out [n_gen][DPS_N_STATES][P]   states after the last step (may be NULL); with a workspace it also carries
 *                                           the state between the step chunks of the work queue
 *   ckpt     out [n_ckpt][n_gen][DPS_N_STATES][P] state *before* step k*ckpt_stride (NULL if ckpt_stride == 0)
 */
int dps_forward(const dps_plan* plan, int64_t P,
                const double* x0, const double* upar, const double* vpar, const double* zsub,
                double* rec, double* x_final, double* ckpt,
                void* workspace, size_t workspace_bytes,   /* dps_forward_workspace_bytes(); NULL / 0 = static launch */
                void* stream);

/*
 * Same as dps_forward for the step range [step_begin, step_end) only, starting from the state `x_in` (the state before
 * step `step_begin`) and leaving the state after step `step_end - 1` in `x_out` (may alias x_in).  `rec` and `ckpt` are
 * the full-run buffers; only the rows of the range are written.  Used to stream the recorder to the host while later
 * steps are still being integrated, and to apply parameter events between ranges
 * (ParamEvent.handle_event, static_models.py:82-95, simulator.py:533-534).
 */
int dps_forward_range(const dps_plan* plan, int64_t P, int32_t step_begin, int32_t step_end,
                      const double* x_in, const double* upar, const double* vpar, const double* zsub,
                      double* rec, double* x_out, double* ckpt,
                      void* workspace, size_t workspace_bytes, void* stream);

/*
 * Discrete adjoint of dps_forward (what torch autograd does for the reference in
 * optimization_lib/ps_optimization.py:212).  Recomputes each checkpoint segment forward, then sweeps it
 * backwards.  Gradients are with respect to the kernel inputs; the host chains them into the torch graph of
 * the prologue, including -- on request -- dL/dZsub for parameters that enter Y_dyn (g_zsub below).
 *
 *   ckpt          [n_ckpt][n_gen][DPS_N_STATES][P]  checkpoints written by dps_forward (ckpt_stride > 0)
 *   grad_rec      [n_rec][n_steps][P]   dL/d rec
 *   grad_x_final  [n_gen][DPS_N_STATES][P] dL/d x_final, or NULL
 *   want          HOST [n_gen*DPS_N_SLOTS] nonzero where a constant-slot gradient is required
 *   g_x0          out [n_gen][DPS_N_STATES][P]
 *   gslot         out [n_want][P]  per-trajectory gradient of every wanted slot, in the order they appear in `want`
 *                 (the host sums over p when the slot really is shared by the batch); may be NULL if nothing is wanted
 *   g_zsub        in/out [n_cfg][z_rows][z_cols][2][P] or NULL: per-trajectory dL/dZsub (rows < n_obs, columns < n_gen),
 *                 ACCUMULATED into a buffer the caller zeroed -- also per trajectory when Zsub is shared by the batch
 *                 (the host then sums over p).  From V = Z I: dZ[r][j] += dV_r conj(I_j) in (re, im) components.
 *   workspace     device scratch of dps_workspace_bytes(plan, P) bytes
 */
int dps_backward(const dps_plan* plan, int64_t P,
                 const double* upar, const double* vpar, const double* zsub, const double* ckpt,
                 const double* grad_rec, const double* grad_x_final,
                 const uint8_t* want, double* g_x0, double* gslot, double* g_zsub,
                 void* workspace, size_t workspace_bytes, void* stream);

/* ---- the identification step on the device (reference: optimization_lib/) ------------------------
 *
 * dps_ident_loss: the reference's default loss with its min/max normalisation and the gradient autograd would return for
 * ``loss.sum().backward()`` (ps_optimization.py:87-101, 178-199, 202, 212):
 *     loss[p] = mean_t sum_c | (target - vmin)/vrange - (rec - vmin)/vrange |          (vmin/vrange NULL: no normalisation)
 *   rec      [C][T][P]   recorder output of dps_forward
 *   target   [C][T][Pt]  Pt = 1 (one target trajectory for the whole batch) or P
 *   vmin, vrange [C][Pt] or both NULL
 *   loss     out [P]
 *   g_rec    out [C][T][P] d(sum_p loss[p])/d rec, or NULL
 *   state    in/out [4] doubles {max_step, last_min_loss, cur_min_loss, cur_min_row} or NULL: cur_min_* receive the minimum
 *            of nan_to_num(loss, 1e5) over the batch and its row (ps_optimization.py:202)
 *   workspace  dps_ident_loss_workspace_bytes(P, T) bytes
 */
size_t dps_ident_loss_workspace_bytes(int64_t P, int32_t T);
int dps_ident_loss(int64_t P, int32_t T, int32_t C, int64_t Pt, const double* rec, const double* target,
                   const double* vmin, const double* vrange, double* loss, double* g_rec, double* state,
                   void* workspace, size_t workspace_bytes, void* stream);

/* dps_bfgs_step: CustomBFGSREALOptimizer.step() (optimizers.py:52-111) for every trajectory at once: BFGS on the inverse
 * Hessian without line search; the first step is scaled to the maximum relative step, every step is clipped to it.
 *   n            optimised parameters per trajectory (1..8)
 *   x, grad      HOST arrays of n device pointers: real part of parameter i / of its gradient for trajectory p at
 *                x[i][p * x_stride[i]] (stride 2 for complex128 [P,1] tensors, 1 for float64); x is updated in place
 *                (``p.data += sk``, :83-85)
 *   hesse        in/out [n][n][P]   last_grad, last_sk  in/out [n][P]
 *   state        [4] doubles as above; state[0] (the current maximum relative step) is read
 *   first_step   1: initialise hesse as :64-70.   have_last 0: no update of hesse yet (:89-92)
 */
int dps_bfgs_step(int64_t P, int32_t n, double* const* x, const int32_t* x_stride, const double* const* grad,
                  const int32_t* grad_stride, double* hesse, double* last_grad, double* last_sk, const double* state,
                  int32_t first_step, int32_t have_last, void* stream);

/* dps_ident_control: ps_optimization.py:246-249 on the device state:
 *     if last_min_loss and last_min_loss < cur_min_loss: max_step *= decay;   last_min_loss = cur_min_loss */
int dps_ident_control(double* state, double decay, void* stream);

/* ---- initialize() on the device ----------------------------------------------------------
 * What PowerSystemSimulation.run() does before the time loop (simulator.py:438-450), for callers that need no autograd
 * graph through it.  Complex arrays are interleaved (re, im) pairs -- torch.complex128 memory -- with the batch index
 * OUTERMOST ([B][n]...); a batch size argument of 1 means "shared by all B rows".  Networks of up to 64 buses (dense
 * workspaces live in shared memory); rc 2 = outside the supported size, the caller keeps its own path.               */
enum dps_bus_type { DPS_BUS_SLACK = 0, DPS_BUS_PV = 1, DPS_BUS_PQ = 2 };          /* load_flow.py:37-60 */

/* raw model parameters of one generator, as the reference's model constructors hold them
 * (synchronous_machine.py:31-84, exciters.py:35-72, governors.py:35-75, stabilizers.py:35-73)                        */
enum dps_raw {
    DPS_RAW_SN = 0, DPS_RAW_SN_SYS, DPS_RAW_XD, DPS_RAW_XQ, DPS_RAW_XDT, DPS_RAW_XQT, DPS_RAW_XDST, DPS_RAW_XQST,
    DPS_RAW_TD0T, DPS_RAW_TQ0T, DPS_RAW_TD0ST, DPS_RAW_TQ0ST, DPS_RAW_H, DPS_RAW_D,
    /* SEXS: lead-lag T_1 (T_a), T_2 (T_b); PT1Limited T (T_e), gain K, limits; voltage set point */
    DPS_RAW_A_T1, DPS_RAW_A_T2, DPS_RAW_A_TE, DPS_RAW_A_K, DPS_RAW_A_EMIN, DPS_RAW_A_EMAX, DPS_RAW_A_VSET,
    /* TGOV1: PT1Limited T (T_1), gain 1/R, limits; lead-lag T_1 (T_2), T_2 (T_3); D_t */
    DPS_RAW_G_T1, DPS_RAW_G_GAIN, DPS_RAW_G_VMIN, DPS_RAW_G_VMAX, DPS_RAW_G_T2, DPS_RAW_G_T3, DPS_RAW_G_DT,
    /* STAB1: washout K_w, T_w; lead-lag 1 (T_1, T_2), lead-lag 2 (T_1, T_2) as stabilizers.py:71-72 wires them; limit */
    DPS_RAW_S_KW, DPS_RAW_S_TW, DPS_RAW_S_LL1_T1, DPS_RAW_S_LL1_T2, DPS_RAW_S_LL2_T1, DPS_RAW_S_LL2_T2, DPS_RAW_S_HLIM,
    DPS_N_RAW
};

/* dps_load_flow: polar Newton-Raphson power flow (load_flow.py:9-179), one thread block per batch row: mismatch,
 * Jacobian, dense solve with partial pivoting and the polar update of one iteration are one kernel launch.  Same unknown
 * ordering (angles of PV then PQ buses, magnitudes of PQ buses), update sequence and stopping rule as the reference:
 * iterate until the batch-global max |mismatch| of the PRE-update iterate is below tol (load_flow.py:147-167) -- a host
 * decision, so this entry point SYNCHRONISES the stream once per iteration (the one exception to "no hidden
 * synchronisation"; it is a once-per-run prologue call).  rc 4 = no convergence after max_iter iterations.
 *   bus_type   HOST [n_bus] dps_bus_type            y_static  [By][n][n] complex       s_soll  [Bs][n] complex
 *   v_init     [Bv][n] complex (flat start / set points)       v_out, s_out  [B][n] complex (voltage, calculated power)
 *   worst      device scratch, one double            iters     HOST out: iterations after the first                   */
int dps_load_flow(int32_t n_bus, const int32_t* bus_type, int64_t B, const double* y_static, int64_t By,
                  const double* s_soll, int64_t Bs, const double* v_init, int64_t Bv, double* v_out, double* s_out,
                  double* worst, double tol, int32_t max_iter, int32_t* iters, void* stream);

/* dps_dynamic_inverse: loads become constant admittances from the load-flow voltage, y = 1 / conj(|v|^2 / s_load)
 * (static_models.py:515-527), are added to the diagonal of y_base (branches, shunts, generator Norton admittances:
 * simulator.py:400-422), and Z0 = inv(Y_dyn) (simulator.py:386) by complex Gauss-Jordan elimination with partial
 * pivoting, one thread block per batch row.
 *   load_bus  HOST [n_load]     s_load [Bl][n_load] complex (per unit)     v [Bv][n] complex
 *   y_load    out [B][n_load] complex    y_diag  out [B][n] complex (diagonal of Y_dyn)    z0  out [B][n][n] complex */
int dps_dynamic_inverse(int32_t n_bus, int64_t B, const double* y_base, int64_t By, int32_t n_load, const int32_t* load_bus,
                        const double* s_load, int64_t Bl, const double* v, int64_t Bv, double* y_load, double* y_diag,
                        double* z0, void* stream);

/* dps_model_init: initial states (canonical dps_state layout), derived constants (dps_slot) and the Norton current at
 * x0 of every (trajectory, generator): synchronous_machine.py:279-323, 256-260; exciters.py:129-138;
 * governors.py:136-145; stabilizers.py:135-145; blocks.py:196-207, 332-342, 426-436.  Any number of buses up to 32767.
 *   gen_bus, ctrl  HOST [n_gen] (bus index, dps_ctrl flags)      v, s  [Bv|Bs][n_bus] complex load-flow voltage / power
 *   raw_u     [n_gen][DPS_N_RAW] batch-uniform raw parameters    raw_vary  HOST [n_gen][DPS_N_RAW]: -1 or row of raw_v
 *   raw_v     [n_rv][B] per-trajectory raw parameters (may be NULL when nothing varies)
 *   x0   out [n_gen][13][B]      slots  out [n_gen][DPS_N_SLOTS][B]      inj  out [n_gen][2][B] (re, im)              */
int dps_model_init(int32_t n_gen, int32_t n_bus, int64_t B, const int32_t* gen_bus, const int32_t* ctrl, const double* v,
                   int64_t Bv, const double* s, int64_t Bs, const double* raw_u, const int32_t* raw_vary, const double* raw_v,
                   double* x0, double* slots, double* inj, void* stream);

/* dps_recorder_to_host: the time slice [first, first + count) (in elements: step * P .. ) of every row of the recorder
 * rec[n_rows][row_elems = T * P] to the same place of a pinned HOST buffer of the same shape, as ONE strided asynchronous
 * copy (the reference keeps its recorder on the host: simulator.py:540-546).  Used to stream finished step ranges to the
 * host while later steps are still being integrated (Pss.run(host_out=...)).                                          */
int dps_recorder_to_host(double* host, const double* rec, int32_t n_rows, int64_t row_elems, int64_t first, int64_t count,
                         void* stream);

/* Measurement helper: runs a dependent-chain-free DFMA loop on every SM and returns the achieved fp64
 * TFLOP/s (2 flops per DFMA) in *tflops.  Synchronises.  Used by bench.py as the fp64 roofline denominator
 * (MEASURED_PEAKS.json has no fp64 entry). */
int dps_fp64_peak_probe(int32_t iters, double* tflops, double* milliseconds);

/* Same measurement with the FP64 tensor-core instruction (mma.sync.m8n8k4.f64, SASS DMMA; 512 flops per warp
 * instruction).  On B200 it shares the FP64 pipe with DFMA; bench.py uses max(DFMA, DMMA) as the roofline peak. */
int dps_fp64_dmma_probe(int32_t iters, double* tflops, double* milliseconds);

/* Test helper: evaluates the kernels' own branch-free math primitives on the device,
 * out[0..4][n] = sin(a), cos(a), a/b, sqrt(|a|), 1/b  (a, b, out: device arrays). */
int dps_math_probe(int64_t n, const double* a, const double* b, double* out, void* stream);

/* Static information about the specialisation a plan selected (for logs / bench "gpu_launches"). */
int dps_plan_info(const dps_plan* plan, int32_t* block_threads, int32_t* sims_per_block, int32_t* regs_per_thread,
                  int32_t* g_template);

const char* dps_last_error(void);
int dps_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DPS_H */
