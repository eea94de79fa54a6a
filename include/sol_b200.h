/* sol_b200.h — C ABI of the B200-native solver-in-the-loop hot path.
 *
 * The reference (tum-pbs/Solver-in-the-Loop) has no FFI of its own: its hot path sits behind
 * Python call signatures (KarmanFlow.step, to_feature, model(...), to_staggered) and PhiFlow's
 * pressure-solver plug-in slot.  Each entry point below names the reference
 * interface it replaces.  Conventions:
 *   - every function returns an int status (SOL_OK = 0); nothing throws across the boundary;
 *     sol_last_error_string() describes the last failure on the calling thread;
 *   - all data pointers are CUDA *device* pointers owned by the caller (PyTorch tensors in the
 *     shipped host layer), except where a parameter is documented as host memory;
 *   - work is enqueued asynchronously on the given cudaStream_t (passed as void*);
 *   - nothing is allocated or freed across the ABI except the opaque plan / unroll handles
 *     (a plan owns a few KB of device-resident masks; an unroll object carves a caller-provided
 *     workspace);
 *   - fp32, contiguous struct-of-arrays fields, x fastest:
 *       vy  [B, Y+1, X]   y-faces  (reference velocity.data[0], karman_train.py:177)
 *       vx  [B, Y, X+1]   x-faces  (reference velocity.data[1], karman_train.py:178)
 *       rho, p, div [B, Y, X]
 *     CNN tensors are NHWC [B, Y, X, C]; CNN parameters are one flat buffer in Keras order
 *     (kernel [5,5,Cin,Cout] then bias [Cout], layer after layer; karman_train.py:101-138).
 */
#ifndef SOL_B200_H
#define SOL_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SOL_ABI_VERSION 1

enum {
    SOL_OK = 0,
    SOL_ERR_INVALID = 1,      /* bad argument */
    SOL_ERR_CUDA = 2,         /* a CUDA runtime call failed */
    SOL_ERR_UNSUPPORTED = 3,  /* shape / model not supported by the kernels */
    SOL_ERR_WORKSPACE = 4     /* caller workspace too small */
};

enum { SOL_BOUNDARY_OPEN = 0, SOL_BOUNDARY_PERIODIC = 1 };
enum { SOL_MODEL_MARS_MOON = 0, SOL_MODEL_MERCURY = 1 };
enum { SOL_ACT_NONE = 0, SOL_ACT_LRELU = 1, SOL_ACT_DLRELU = 2 };

typedef struct sol_plan sol_plan;
typedef struct sol_unroll sol_unroll;

int sol_abi_version(void);
const char* sol_last_error_string(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
unsigned long long sol_launch_count(void);

/* ---- plan: scene geometry -------------------------------------------------------------------
 * Replaces KarmanFlow.__init__ + Domain/Fluid construction (karman_train.py:166-171, 363-372).
 * Host arrays (copied to the device once):
 *   solid     [Y*X]     uint8, 1 = obstacle cell           (Obstacle(Sphere([50,50],10)))
 *   inflow    [Y*X]     float, density inflow rate per cell (Inflow(box[5:10,25:75])), may be NULL
 *   bc_mask_y [(Y+1)*X] float, velBCyMask, may be NULL     (karman_train.py:366-372)
 *   bc_val_y  [(Y+1)*X] float, velBCy
 * dx = cell size in physical units (len/res); boundary = SOL_BOUNDARY_OPEN (karman) or
 * SOL_BOUNDARY_PERIODIC (burgers; solid/inflow/bc must be NULL). */
int sol_plan_create(int Y, int X, int B_max, float dx, int boundary,
                    const unsigned char* solid, const float* inflow,
                    const float* bc_mask_y, const float* bc_val_y, sol_plan** out);
int sol_plan_destroy(sol_plan* plan);
/* CG controls: stop when max|r| < max(tol_abs, tol_rel*max|rhs|) per simulation, or after max_it
 * iterations (reference SparseCG: tol_abs=1e-5, tol_rel=0, max_it=2000).  cluster = CTAs per
 * simulation (1,2,4,8; 0 = auto). */
int sol_plan_set_cg(sol_plan* plan, float tol_abs, float tol_rel, int max_it, int cluster);
/* Tuning knobs that never change results beyond fp32 round-off:
 *   "cg_rows"    rows of the grid each CG thread keeps in registers (0 = auto, 2/4/8/16)
 *   "cg_precond" 1 (default) = multigrid-preconditioned CG where the grid supports it (same stop
 *                rule, ~15x fewer iterations), 0 = the reference's unpreconditioned recurrences
 *   "direct_solve" 1 (default) = direct projection where the scene supports it (OPEN plans of 128x64 / 64x32 cells, cluster <= 1):
 *                fast Poisson solve (sine transforms) + capacitance-matrix correction for the obstacle rows, precomputed per plan
 *                on first use; exact up to fp32 round-off (what the reference's NumPy path does with a sparse direct solver,
 *                karman_apply.py:39), the CG controls above do not apply and iteration counters read 0.  0 = iterative solvers */
int sol_plan_set_option(sol_plan* plan, const char* name, int value);
/* Read-only plan facts for measurement scripts: "direct_active" (1 when the direct projection is what sol_project runs for
 * batch 1), "direct_rows" (rows of the pressure operator the obstacle changes, padded to 32: the k of the capacitance matrix),
 * "sm_count". */
int sol_plan_query(sol_plan* plan, const char* name, int* value);
/* Process-wide knobs:
 *   "conv_path" 0 = auto (= 2), 1 = fp32 SIMT kernels, 2 = tcgen05 tensor-core kernels with block-scaled 3xFP16 operand
 *         splitting (fp32-accurate), 3 = tcgen05 with 3xTF32 splitting (the round-1 kernel, twice the operand traffic)
 *   "conv_variant" (tuning) accumulator layout of the 3xFP16 kernel: 0 = two merged sets (default), 1 = two sets, 2 = one set
 *   "wgrad_path" 0 = auto (= 2), 1 = per-step fp32 SIMT weight gradients, 2 = deferred tcgen05 weight-gradient GEMM with
 *         block-scaled 3xFP16 operands, 3 = deferred tcgen05 GEMM with 3xTF32 operands (the round-1 kernel)
 *   "fuse_small" 1 = the correction-gradient scaling (adjoint of "velocity + correction") is folded into the
 *         diffusion adjoint of the following step, 0 (default: measured faster) = separate kernel
 *   "fuse_solver_io" 1 (default) = to_feature and its adjoint are folded into the projection kernel (2 launches fewer per
 *         step) where the solver variant supports it, 0 = separate kernels
 *   "wgrad_overlap" 1 (default) = the deferred weight-gradient GEMMs of already finished steps run on a side stream
 *         while an adjoint pressure solve keeps only B SMs busy, 0 = all of them after the adjoint sweep
 *   "wgrad_window_us" (tuning) time budget of one such window at 128x64, default 110
 *   "pdl" 1 (default) = kernels are launched with programmatic stream serialization (the prologue of a kernel
 *         overlaps the tail of its predecessor on the stream), 0 = plain stream order
 *   "wgrad_bg_ctas" / "wgrad_bg_chunk" (tuning) with the direct projection there is no solve window: the weight-gradient items of
 *         finished steps then run as persistent background launches of at most this many CTAs (default 48, 0 = off; one launch
 *         per layer per chunk of steps, default 4) on the SMs the adjoint conv chain does not need
 *   "deterministic" 1 = the reductions the engine has a choice about are ordered instead of atomic: the weight gradients of the
 *         first / last layer go through private CTA slots, the per-step losses through per-CTA partials, both summed in CTA
 *         order; bit-identical from run to run EXCEPT for the scatter-add of the advection adjoint, whose floating-point
 *         atomics (sol_advect_bwd) remain order-dependent at round-off level.  0 (default) = atomics
 *   "fuse_stencil" 1 (default) = viscosity + BC and the three advections of a step are ONE launch with the stencil halo staged in
 *         shared memory (OPEN plans), 0 = one kernel per stage (global gathers)
 *   "thin_path" 0 (default) = first / last conv layers (Cin <= 4 -> 32, 32 -> Cout <= 4) on the row-pair kernels of
 *         sol_conv_thin.cu, 1 = the first-generation one-pixel-per-thread kernels (validation)
 *   "nvtx" 1 = NVTX ranges around the stages of the unrolled sweeps (for nsys / ncu --nvtx; host-side, default 0)
 *   "tc_base_offset_mode" (debug) how the UMMA shared-memory descriptors encode unaligned starts */
int sol_set_option(const char* name, int value);

/* ---- stage entry points (one reference op each) --------------------------------------------- */

/* diffuse(CenteredGrid(v), alpha) + velocity BC, alpha_b = dt*res^2/re_b (karman_train.py:175-181) */
int sol_diffuse_bc(sol_plan* plan, void* stream, int B, const float* re, float dt, float res,
                   const float* vy, const float* vx, float* vy_out, float* vx_out);
/* adjoint of sol_diffuse_bc w.r.t. (vy, vx) */
int sol_diffuse_bc_bwd(sol_plan* plan, void* stream, int B, const float* re, float dt, float res,
                       const float* g_vy_out, const float* g_vx_out, float* g_vy, float* g_vx);

/* sol_diffuse_bc followed by sol_advect in ONE launch (OPEN plans): the viscosity + BC result (vy1, vx1: also the stash of the
 * advection adjoint) is staged in shared memory with its halo and advected from there; identical results. rho/rho_out may be NULL. */
int sol_diffuse_advect(sol_plan* plan, void* stream, int B, const float* re, float dt, float res, const float* vy, const float* vx,
                       const float* rho, float* vy1, float* vx1, float* vy2, float* vx2, float* rho_out);

/* advect.semi_lagrangian(density|velocity, velocity, dt) + Inflow effect
 * (IncompressibleFlow.step via karman_train.py:185).  rho/rho_out may be NULL. */
int sol_advect(sol_plan* plan, void* stream, int B, float dt, const float* vy, const float* vx,
               const float* rho, float* vy_out, float* vx_out, float* rho_out);
/* adjoint of the velocity self-advection; g_vy/g_vx are OVERWRITTEN (zeroed then scattered into) */
int sol_advect_bwd(sol_plan* plan, void* stream, int B, float dt, const float* vy, const float* vx,
                   const float* g_vy_out, const float* g_vx_out, float* g_vy, float* g_vx);

/* PoissonSolver.solve(divergence, fluiddomain) — the pressure-solver plug-in slot the reference
 * scripts name (karman_train.py:51,167-168).  div, p: [B,Y,X]; iters: int [B] or NULL. */
int sol_pressure_solve(sol_plan* plan, void* stream, int B, const float* div, float* p, int* iters);
/* divergence_free(velocity, domain, obstacles): hard-BC mask, divergence, CG solve, gradient
 * subtract, fused in one persistent kernel.  p_out / iters may be NULL.  Self-adjoint: the same
 * call applied to an upstream gradient is the adjoint. */
int sol_project(sol_plan* plan, void* stream, int B, const float* vy, const float* vx,
                float* vy_out, float* vx_out, float* p_out, int* iters);
/* stand-alone masked divergence (FluidDomain.with_hard_boundary_conditions + divergence) */
int sol_divergence(sol_plan* plan, void* stream, int B, const float* vy, const float* vx, float* div);

/* one KarmanFlow.step (karman_train.py:173-185): diffuse+BC -> advect -> project.
 * vy1/vx1 (post-BC velocity, the advection adjoint's input) may be NULL when no adjoint is needed;
 * rho_in/rho_out, p_out, iters may be NULL. */
int sol_step_fwd(sol_plan* plan, void* stream, int B, const float* re, float dt, float res,
                 const float* rho_in, const float* vy_in, const float* vx_in,
                 float* rho_out, float* vy_out, float* vx_out, float* p_out,
                 float* vy1, float* vx1, float* scratch_vy, float* scratch_vx, int* iters);
/* adjoint of sol_step_fwd w.r.t. the input velocity; scratch_* are [B,..] face arrays */
int sol_step_bwd(sol_plan* plan, void* stream, int B, const float* re, float dt, float res,
                 const float* vy1, const float* vx1, const float* g_vy_out, const float* g_vx_out,
                 float* g_vy_in, float* g_vx_in, float* scratch_vy, float* scratch_vx, int* iters);

/* Burgers.step / BurgersTest.step_with_f (burgers_train.py:178-187): periodic advect ->
 * diffuse -> + dt*f.  diff_kernel_{y,x}: device real-space circular kernels of the periodic
 * diffusion operator on the [Y+1,X] / [Y,X+1] component arrays (NULL: explicit 5-point with
 * amount = viscosity*dt).  fy/fx may be NULL. */
int sol_burgers_step(sol_plan* plan, void* stream, int B, float dt, float viscosity,
                     const float* diff_kernel_y, const float* diff_kernel_x,
                     const float* vy, const float* vx, const float* fy, const float* fx,
                     float* vy_out, float* vx_out, float* scratch_vy, float* scratch_vx);
int sol_burgers_step_bwd(sol_plan* plan, void* stream, int B, float dt, float viscosity,
                         const float* diff_kernel_y, const float* diff_kernel_x,
                         const float* vy, const float* vx, const float* g_vy_out, const float* g_vx_out,
                         float* g_vy, float* g_vx, float* scratch_vy, float* scratch_vx);

/* ---- correction network (keras Conv2D 5x5 'same', karman_train.py:101-138) ------------------ */

/* out = act(conv5x5(in; w) + bias + addend); in [B,Y,X,Cin], w [5,5,Cin,Cout], out [B,Y,X,Cout].
 * act: SOL_ACT_NONE | SOL_ACT_LRELU (slope) | SOL_ACT_DLRELU (multiply by 1 or slope by sign of
 * ref).  bias/addend/ref may be NULL. */
int sol_conv5x5(void* stream, int B, int Y, int X, int Cin, int Cout, const float* in, const float* w,
                const float* bias, const float* addend, const float* ref, int act, float slope, float* out);
/* Tensor-core path of the 32->32 layers (same keras Conv2D, karman_train.py:107-133): the fp32 weights
 * are split ONCE per optimiser step into their tf32 hi/lo operand layout (sol_conv5x5_split_floats()
 * floats, 16-byte aligned) and reused by every unrolled step; sol_conv5x5() splits on every call.
 * weights_settled != 0 is the caller's guarantee that wsplit was complete before the PREVIOUS kernel of this
 * stream was launched (e.g. split once per optimiser step): the kernel then prefetches the weights while its
 * predecessor drains (programmatic dependent launch).  Pass 0 when in doubt. */
size_t sol_conv5x5_split_floats(void);
int sol_conv5x5_split_weights(void* stream, const float* w, float* wsplit);
int sol_conv5x5_c32_presplit(void* stream, int B, int Y, int X, const float* in, const float* wsplit, const float* bias,
                             const float* addend, const float* ref, int act, float slope, float* out, int weights_settled);
/* wT[5,5,Cout,Cin] = flip+transpose of w[5,5,Cin,Cout]: conv5x5(g_out; wT) is the data gradient */
int sol_conv5x5_flip_weights(void* stream, int Cin, int Cout, const float* w, float* wT);
/* dW[5,5,Cin,Cout] (+)= in (x) g_out, db[Cout] (+)= sum g_out.  partials: workspace of
 * sol_conv5x5_wgrad_workspace(Cin,Cout) floats (may be NULL for thin layers). */
size_t sol_conv5x5_wgrad_workspace(int Cin, int Cout);
int sol_conv5x5_wgrad(void* stream, int B, int Y, int X, int Cin, int Cout, const float* in, const float* g_out,
                      float* dW, float* db, int accumulate, float* partials);

size_t sol_model_param_count(int model, int cin0);

/* to_feature(state, Re)/sigma (karman_train.py:77-86,416-420): feat[B,Y,X,3] */
int sol_to_feature(sol_plan* plan, void* stream, int B, const float* vy, const float* vx, const float* re,
                   float sig_vy, float sig_vx, float sig_re, float* feat);

/* ---- the unrolled training iteration (karman_train.py:393-457, sess.run at :502) ------------ */
typedef struct {
    int model;          /* SOL_MODEL_* */
    int cin0;           /* feature channels: 3 = (vy, vx, Re) for karman */
    int msteps;         /* unrolled solver steps */
    int B;              /* simulations in this batch (<= plan B_max) */
    float dt;
    float res;          /* reference resolution (X) used in alpha = dt*res^2/Re */
    float sig_vy, sig_vx, sig_ext;   /* dataStats std: velocity comps and Re (karman_train.py:234-255) */
    int with_density;   /* advect the (forward-only) marker density too */
    int use_graph;      /* capture fwd+bwd into a CUDA graph and replay it */
} sol_unroll_cfg;

size_t sol_unroll_workspace_bytes(const sol_plan* plan, const sol_unroll_cfg* cfg);
int sol_unroll_create(sol_plan* plan, const sol_unroll_cfg* cfg, void* workspace, size_t workspace_bytes,
                      sol_unroll** out);
int sol_unroll_destroy(sol_unroll* u);
/* forward: msteps x (step -> CNN -> add correction), per-step l2 losses (tf.nn.l2_loss, :428-436).
 * gt_vy [m,B,Y+1,X], gt_vx [m,B,Y,X+1] (NULL: no loss); loss_steps float[m];
 * pred_* optional outputs [m,B,...] of the corrected states. */
int sol_unroll_forward(sol_unroll* u, void* stream, const float* weights, const float* re,
                       const float* rho0, const float* vy0, const float* vx0,
                       const float* gt_vy, const float* gt_vx, float* loss_steps,
                       float* pred_vy, float* pred_vx, float* pred_rho);
/* forward-only rollout — the loop of karman_apply.py:138-151 / burgers_apply.py:129-151: nsteps x (step -> correction net -> add),
 * every corrected frame written to pred_vy [nsteps,B,Y+1,X] / pred_vx [nsteps,B,Y,X+1] (/ pred_rho [nsteps,B,Y,X] with rho0 and
 * cfg.with_density).  nsteps is independent of cfg.msteps (no adjoint stash is kept; sol_unroll_backward is invalid afterwards).
 * Burgers unrolls: the force arrays registered with sol_unroll_set_burgers must hold nsteps frames. */
int sol_unroll_rollout(sol_unroll* u, void* stream, const float* weights, const float* re,
                       const float* rho0, const float* vy0, const float* vx0, int nsteps,
                       float* pred_vy, float* pred_vx, float* pred_rho);
/* adjoint sweep of the last forward: grad_weights (flat, Keras order) = d(sum_i loss_i / m)/d(weights);
 * g_vy0/g_vx0 (optional) receive the gradient w.r.t. the initial velocity. */
int sol_unroll_backward(sol_unroll* u, void* stream, const float* weights, float* grad_weights,
                        float* g_vy0, float* g_vx0);
/* forward + backward in one call (CUDA-graph replay when cfg.use_graph) */
int sol_unroll_train_iter(sol_unroll* u, void* stream, const float* weights, const float* re,
                          const float* rho0, const float* vy0, const float* vx0,
                          const float* gt_vy, const float* gt_vx, float* loss_steps, float* grad_weights);
/* Burgers scene of an unroll created on a PERIODIC plan (burgers/burgers_train.py:379-437): every unrolled step is
 * BurgersTest.step_with_f (:178-187, advect -> diffuse(viscosity*dt) -> + dt*f_i) followed by the correction net on
 * to_feature([v],[f_i])/[std_v, std_f] (:75-82, 398-415; cfg.cin0 = 4) or to_feature_noforce (:84-90; cfg.cin0 = 2,
 * f_* = NULL means step() without forcing).  diff_kernel_{y,x} as in sol_burgers_step (NULL: explicit 5-point);
 * f_vy [m,B,Y+1,X], f_vx [m,B,Y,X+1] device arrays that must stay valid while the unroll is used.  cfg.sig_vy/sig_vx
 * are the velocity std (features, output scaling and loss), sig_fy/sig_fx the force std; `re` of the sol_unroll_*
 * calls is ignored (may be NULL).  Must be called before the first forward. */
int sol_unroll_set_burgers(sol_unroll* u, float viscosity, const float* diff_kernel_y, const float* diff_kernel_x,
                           const float* f_vy, const float* f_vx, float sig_fy, float sig_fx);
/* device pointer to int[2*msteps*B] CG iteration counts: [0..m*B) forward, [m*B..2mB) adjoint */
int sol_unroll_cg_iters(sol_unroll* u, const int** dev_iters, int* count);

/* tf.compat.v1.train.AdamOptimizer update (karman_train.py:449-457) on flat fp32 buffers:
 * lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t*m/(sqrt(v)+eps).  grad_scale multiplies g first. */
int sol_adam_tf1(void* stream, size_t n, float* theta, const float* grad, float* m, float* v,
                 int t, float lr, float beta1, float beta2, float eps, float grad_scale);

#ifdef __cplusplus
}
#endif
#endif /* SOL_B200_H */
