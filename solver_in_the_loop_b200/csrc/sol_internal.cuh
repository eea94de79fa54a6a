// Internal declarations shared by the translation units of libsol_b200.so.
#pragma once
#include <utility>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <vector>

#include "../../include/sol_b200.h"

namespace sol {

extern thread_local char g_err[512];
extern std::atomic<unsigned long long> g_launches;

inline int fail(int code, const char* fmt, const char* a = "", const char* b = "") {
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}

#define SOL_CUDA(call)                                                                           \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            snprintf(sol::g_err, sizeof(sol::g_err), "%s:%d: %s -> %s", __FILE__, __LINE__, #call, \
                     cudaGetErrorString(e__));                                                   \
            return SOL_ERR_CUDA;                                                                 \
        }                                                                                        \
    } while (0)

#define SOL_CHECK(cond, msg)                                                             \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            snprintf(sol::g_err, sizeof(sol::g_err), "%s:%d: %s", __FILE__, __LINE__, msg); \
            return SOL_ERR_INVALID;                                                      \
        }                                                                                \
    } while (0)

#define SOL_TRY(call)            \
    do {                         \
        int rc__ = (call);       \
        if (rc__ != SOL_OK) return rc__; \
    } while (0)

// count + check a kernel launch
#define SOL_LAUNCHED()                                   \
    do {                                                 \
        sol::g_launches.fetch_add(1, std::memory_order_relaxed); \
        SOL_CUDA(cudaGetLastError());                    \
    } while (0)

// Programmatic dependent launch: every kernel of this library is launched with programmatic stream
// serialization, so its CTAs may become resident while the previous kernel of the stream is still
// draining.  A kernel must therefore execute pdl_wait() before its first access to global memory that
// another kernel reads or writes (pdl_sync() at the top, or later when there is launch-independent
// prologue work), and triggers its own dependents only after that wait, which makes completion
// transitive along the stream.  With the attribute off (option "pdl" = 0) both are no-ops.
extern int g_pdl;

// Diagnostics (sol_debug_chain_trace, scripts/chain_trace.py): with a trace buffer installed, thread 0 of CTA 0 of EVERY kernel
// stamps %globaltimer right after its griddepcontrol.wait, i.e. at the moment its predecessor in the stream has completed.
// Kernels of one stream form a serial chain, so the difference of consecutive stamps is the cost of a kernel inside the
// replayed graph (programmatic overlap included) — what a serialising profiler cannot show.  The control block lives in
// __constant__ memory, one copy per translation unit (no relocatable device code): SOL_TRACE_TU() registers the copy's setter.
struct TraceCtl { unsigned long long* buf; unsigned int* count; unsigned int cap; };
void trace_register(void (*setter)(const TraceCtl&));
extern bool g_trace_names;                       // host side: record the kernel of every launch (same order as the stamps)
void trace_record_launch(const void* kern);

template <typename... P, typename... A>
inline cudaError_t launch_kernel(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
    if (g_trace_names) trace_record_launch((const void*)kern);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = g_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<A>(args)...);
}
#ifdef __CUDACC__
static __constant__ TraceCtl c_trace;
#define SOL_TRACE_TU()                                                                                                        \
    namespace {                                                                                                               \
    struct TraceTu {                                                                                                          \
        TraceTu() { sol::trace_register([](const sol::TraceCtl& c) { cudaMemcpyToSymbol(sol::c_trace, &c, sizeof(c)); }); }  \
    } trace_tu__;                                                                                                             \
    }
__device__ __forceinline__ void trace_stamp() {
    if (c_trace.buf && (threadIdx.x | threadIdx.y | threadIdx.z | blockIdx.x | blockIdx.y | blockIdx.z) == 0) {
        const unsigned int s = atomicAdd(c_trace.count, 1u);
        if (s < c_trace.cap) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            c_trace.buf[s] = t;
        }
    }
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); trace_stamp(); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() { pdl_wait(); pdl_trigger(); }
#endif

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace sol

constexpr int MG_MAX_LEVELS = 8;

// geometric multigrid hierarchy of the pressure operator (built once per plan, see sol_cg_mg.cu)
struct sol_mg {
    bool valid = false;
    int nlev = 0;                         // levels including the fine grid (0) and the coarsest
    int LY[MG_MAX_LEVELS] = {0}, LX[MG_MAX_LEVELS] = {0};
    int coff[MG_MAX_LEVELS] = {0};        // offsets of levels 1..nlev-2 in dinv / diag
    float* dinv = nullptr;                // device: -omega/diag on fluid cells, 0 on solid
    float* diag = nullptr;                // device
    float* cinv = nullptr;                // device: dense inverse on the coarsest level
    float omega = 0.8f;
};

// direct pressure solver (sol_direct.cu): transform matrices + capacitance correction, precomputed per plan
struct sol_direct {
    bool tried = false, valid = false;
    int k = 0, kp = 0;                    // rows of A the obstacle changes (padded to a multiple of 32)
    float *Sy = nullptr, *Sx = nullptr, *ilam = nullptr;
    int* rt_col = nullptr; float* rt_val = nullptr;
    float* Wt = nullptr;                  // [kp][N]: the capacitance-corrected basis (W M), transposed
    float* p0 = nullptr;                  // scratch: [B_max][N] obstacle-free solution
    float* zbuf = nullptr;                // scratch: [B_max][N] (streamed kernel of the large grids)
};

struct sol_plan {
    float* cg_any_scratch = nullptr;   // generic-grid CG fallback (sol_cg.cu): [B_max][4][Y*X], allocated on first use
    int Y = 0, X = 0, B_max = 0;
    float dx = 1.f;
    int boundary = SOL_BOUNDARY_OPEN;
    int sm_count = 148;
    // device-resident constants (shared by the whole batch)
    unsigned char* active = nullptr;   // [Y*X] 1 = fluid cell
    float* diag = nullptr;             // [Y*X] #accessible neighbours, >= 1
    float* face_my = nullptr;          // [(Y+1)*X] hard-BC face mask
    float* face_mx = nullptr;          // [Y*(X+1)]
    float* inflow = nullptr;           // [Y*X] or null
    float* bc_mask_y = nullptr;        // [(Y+1)*X] or null
    float* bc_val_y = nullptr;
    // CG controls
    float tol_abs = 1e-5f, tol_rel = 0.f;
    int max_it = 2000;
    int cluster = 0;
    int cg_rows = 0;   // rows per thread in the CG kernel (0 = auto)
    int cg_precond = 1; // 1 = multigrid-preconditioned CG when the grid supports it, 0 = plain CG (reference recurrences)
    sol_mg mg;
    int mg_variant = 0; // 0 auto (compile-time hierarchy when Y == 2X), 2 = generic run-time hierarchy kernel
    int direct_solve = 1; // 1 = direct projection (fast Poisson + capacitance correction) when the scene supports it
    sol_direct dir;
    std::vector<unsigned char> h_active;   // host copies of the masks for the (lazy) direct-solver precomputation
    std::vector<float> h_diag;
    size_t NY() const { return (size_t)(Y + 1) * X; }
    size_t NX() const { return (size_t)Y * (X + 1); }
    size_t NC() const { return (size_t)Y * X; }
};

namespace sol {

// ---- stencil stage launchers (sol_stencil.cu) ----
int launch_diffuse_bc(const sol_plan* p, cudaStream_t st, int B, const float* re, float dt, float res,
                      const float* vy, const float* vx, float* vy_out, float* vx_out);
int launch_diffuse_bc_bwd(const sol_plan* p, cudaStream_t st, int B, const float* re, float dt, float res,
                          const float* gy, const float* gx, float* gy_in, float* gx_in, const float* add_y, const float* add_x,
                          float* g_corr = nullptr, float sy = 1.0f, float sx = 1.0f, float* zero_y = nullptr, float* zero_x = nullptr);
int launch_advect(const sol_plan* p, cudaStream_t st, int B, float dt, const float* vy, const float* vx, const float* rho,
                  float* vy_out, float* vx_out, float* rho_out);
// gy / gx are scatter targets: zeroed here by two memsets unless the caller guarantees they already are (targets_are_zero)
int launch_advect_bwd(const sol_plan* p, cudaStream_t st, int B, float dt, const float* vy, const float* vx,
                      const float* gy_out, const float* gx_out, float* gy, float* gx, bool targets_are_zero = false);
// fused diffuse+BC -> advection with the stencil halo staged in shared memory (sol_stencil_fused.cu; OPEN plans)
extern int g_fuse_stencil;
int launch_diffuse_advect(const sol_plan* p, cudaStream_t st, int B, const float* re, float dt, float res, const float* vy, const float* vx,
                          const float* rho, float* vy1, float* vx1, float* vy2, float* vx2, float* rho_out);
int launch_divergence(const sol_plan* p, cudaStream_t st, int B, const float* vy, const float* vx, float* div);
int launch_to_feature(const sol_plan* p, cudaStream_t st, int B, const float* vy, const float* vx, const float* re,
                      float sy, float sx, float sr, float* feat);
// Burgers features: [vy, vx (, fy, fx)][:Y,:X]/sigma (burgers_train.py:75-82, 398-415); cfeat = 2 (--noforce) or 4
int launch_to_feature_burgers(const sol_plan* p, cudaStream_t st, int B, const float* vy, const float* vx, const float* fy, const float* fx,
                              float sy, float sx, float sfy, float sfx, int cfeat, float* feat);
// G_out = G + add (add_* may be NULL); optional fused corr_bwd output g_corr = sigma * G_out[:Y,:X]
int launch_add_faces(const sol_plan* p, cudaStream_t st, int B, const float* gy, const float* gx, const float* add_y, const float* add_x,
                     float* gy_out, float* gx_out, float* g_corr = nullptr, float sy = 1.0f, float sx = 1.0f);
// v_out = v + sigma*corr (zero on the far row/col); optional loss + loss gradient
int launch_correct_loss(const sol_plan* p, cudaStream_t st, int B, const float* vy, const float* vx, const float* corr,
                        float sy, float sx, const float* gt_vy, const float* gt_vx, float inv_m,
                        float* vy_out, float* vx_out, float* gl_vy, float* gl_vx, float* loss, float* loss_part = nullptr);
// deterministic loss: loss[i] = sum of the per-CTA partials loss_part[i * stride + 0 .. nparts) in CTA order, for i < msteps
int launch_loss_finalize(cudaStream_t st, int msteps, const float* loss_part, int stride, float* loss);
int correct_loss_grid(const sol_plan* p, int B);
// g_corr[B,Y,X,2] = sigma * G[:Y,:X]
int launch_corr_bwd(const sol_plan* p, cudaStream_t st, int B, const float* Gy, const float* Gx, float sy, float sx, float* g_corr);
// G3 = G + g_feat[...,0:2]/sigma (padded)
int launch_feat_bwd(const sol_plan* p, cudaStream_t st, int B, const float* Gy, const float* Gx, const float* g_feat, int cfeat,
                    float sy, float sx, float* Gy_out, float* Gx_out);
int launch_burgers_diffuse(const sol_plan* p, cudaStream_t st, int B, float amount, const float* ky, const float* kx,
                           const float* vy, const float* vx, const float* fy, const float* fx, float dtf,
                           float* vy_out, float* vx_out);
int launch_adam(cudaStream_t st, size_t n, float* theta, const float* g, float* m, float* v, float lr_t, float b1, float b2,
                float eps, float gscale);

// ---- pressure projection (sol_cg.cu) ----
// Optional work fused into the projection kernel (only the compile-time-hierarchy multigrid kernel implements it,
// see cg_fuses()): the CNN feature tensor of the projected velocity (to_feature, karman_train.py:77-86,416-420) as an
// extra output, and the feature gradient added to the incoming velocity gradient (adjoint of the same op) as input.
struct CgFuse {
    float* feat_out = nullptr;          // [B,Y,X,cfeat]: (vy_out[j,i]/sig_y, vx_out[j,i]/sig_x, re[b]/sig_r)
    const float* re = nullptr;
    float isy = 0.f, isx = 0.f, isr = 0.f;
    const float* gfeat_in = nullptr;    // [B,Y,X,cfeat]: vy_in[j,i] += gfeat[...,0]*isy (j < Y), vx_in[j,i] += gfeat[...,1]*isx (i < X)
    int cfeat = 3;
};
bool cg_fuses(const sol_plan* p, int B);     // true when launch_cg(p, .., B, ..) accepts a CgFuse
int launch_cg(const sol_plan* p, cudaStream_t st, int B, int mode, const float* rhs, float* p_out, const float* vy,
              const float* vx, float* vy_out, float* vx_out, int* iters, const CgFuse* fuse = nullptr);

// ---- direct projection (sol_direct.cu) ----
bool direct_supported(const sol_plan* p);
int direct_build(sol_plan* p);          // host precomputation + upload (synchronous; never inside a stream capture)
void direct_free(sol_plan* p);
bool direct_active(const sol_plan* p);  // option on and precomputation available (builds lazily)
bool direct_for_batch(const sol_plan* p, int B);   // ... and the batch is small enough for it to be the faster solver
int launch_direct(const sol_plan* p, cudaStream_t st, int B, int mode, const float* rhs, float* p_out, const float* vy, const float* vx,
                  float* vy_out, float* vx_out, int* iters, const CgFuse* fuse = nullptr);

// ---- multigrid-preconditioned CG (sol_cg_mg.cu) ----
bool mg_supported(const sol_plan* p);
int launch_cg_mg(const sol_plan* p, cudaStream_t st, int B, int mode, const float* rhs, float* p_out, const float* vy,
                 const float* vx, float* vy_out, float* vx_out, int* iters, const CgFuse* fuse = nullptr);
bool mg3_selected(const sol_plan* p);

// ---- convolutions (sol_conv.cu) ----
int launch_conv5x5(cudaStream_t st, int B, int Y, int X, int Cin, int Cout, const float* in, const float* w, const float* bias,
                   const float* addend, const float* ref, int act, float slope, float* out, unsigned int* amax_out = nullptr,
                   bool weights_ready = false);
// first / last layers (sol_conv_thin.cu): SOL_ERR_UNSUPPORTED for channel counts it does not cover.  weights_ready: the weights
// were complete before the PREVIOUS kernel of the stream started, so they may be fetched before the programmatic wait.
int launch_conv5x5_thin(cudaStream_t st, int B, int Y, int X, int Cin, int Cout, const float* in, const float* w, const float* bias,
                        const float* addend, const float* ref, int act, float slope, float* out, unsigned int* amax_out, bool weights_ready);
extern int g_thin_path;
void set_thin_cap_mask(int m);
int launch_flip_weights(cudaStream_t st, int Cin, int Cout, const float* w, float* wT);
size_t wgrad_workspace_floats(int Cin, int Cout);
// the ten 32 -> 32 layers at once: layer l sums nctas10[l] slots of partials + l * part_stride into out + l * (25*32*32 + 32)
int launch_wgrad_finalize_multi(cudaStream_t st, const int* nctas10, const float* partials, size_t part_stride, float* out);
int launch_wgrad(cudaStream_t st, int B, int Y, int X, int Cin, int Cout, const float* in, const float* g_out, float* dW, float* db,
                 int accumulate, float* partials, bool finalize);

int launch_wgrad_thin_multi(cudaStream_t st, int steps, int B, int Y, int X, int Cin, int Cout, const float* in, size_t in_step_stride,
                            const float* g, size_t g_step_stride, float* dW, float* db, int max_ctas = 0, float* part = nullptr);
size_t wgrad_thin_part_floats();     // scratch of the deterministic mode (private CTA slots)
extern int g_deterministic;          // option "deterministic": ordered reductions instead of floating-point atomics (thin wgrads, loss)
int launch_wgrad_finalize_n(cudaStream_t st, int nctas, const float* partials, float* dW, float* db, int accumulate);

// ---- deferred tensor-core weight gradient (sol_wgrad_tc.cu) ----
extern int g_wgrad_path;            // 1 SIMT per step, 2 tcgen05 3xFP16 deferred over the whole sweep (default; 0 = auto = 2), 3 tcgen05 3xTF32 deferred
int launch_wgrad_c32_tc(cudaStream_t st, int sm_count, int steps, int B, int Y, int X, const float* in, size_t in_step_stride,
                        const float* g, size_t g_step_stride, float* part, int* nctas_out, int accumulate = 0);
int launch_colsum32(cudaStream_t st, const float* g, size_t npix, float* db);
// 3xFP16 form (sol_wgrad_h.cu): needs max|in|, max|g| of the tensors (device slots, bit patterns) for its power-of-two scales
int launch_wgrad_c32_h(cudaStream_t st, int sm_count, int steps, int B, int Y, int X, const float* in, size_t in_step_stride,
                       const float* g, size_t g_step_stride, const unsigned int* amax_in, const unsigned int* amax_g, float* part,
                       int* nctas_out, int accumulate = 0);
int launch_amax(cudaStream_t st, const float* x, size_t n, unsigned int* slot);      // slot = max(slot, max|x|)

// ---- tensor-core convolutions (sol_conv_h.cu: 3xFP16, default; sol_conv_tc.cu: 3xTF32) ----
extern int g_conv_path;             // 1 SIMT fp32, 2 tcgen05 3xFP16 (default; option value 0 = auto = 2), 3 tcgen05 3xTF32
extern int g_conv_variant;          // accumulator layout of the 3xFP16 kernel (tuning)
extern int g_tc_base_offset_mode;
bool conv_path_is_tc();
size_t tc_weights_floats();         // floats of one layer's pre-split weights (either layout fits)
size_t h_weights_floats();
int launch_prep_tc_weights(cudaStream_t st, const float* w, float* wprep);
int launch_prep_h_weights(cudaStream_t st, const float* w, float* wsplit, int nlayers = 1, size_t w_stride = 0, size_t split_stride = 0);
int launch_split_weights(cudaStream_t st, const float* w, float* wsplit);      // layout of the selected path
// `nlayers` equally spaced 32->32 weight tensors (w + l*w_stride) -> wsplit + l*tc_weights_floats(): one launch on the 3xFP16 path
int launch_split_weights_multi(cudaStream_t st, const float* w, size_t w_stride, float* wsplit, int nlayers);
int launch_conv5x5_tc(cudaStream_t st, int B, int Y, int X, const float* in, const float* wprep, const float* bias,
                      const float* addend, const float* ref, int act, float slope, float* out, bool weights_ready);
int launch_conv5x5_h(cudaStream_t st, int B, int Y, int X, const float* in, const float* wsplit, const float* bias,
                     const float* addend, const float* ref, int act, float slope, float* out, bool weights_ready, unsigned int* amax_out = nullptr);
int launch_conv5x5_c32_presplit(cudaStream_t st, int B, int Y, int X, const float* in, const float* wsplit, const float* bias,
                                const float* addend, const float* ref, int act, float slope, float* out, bool weights_ready,
                                unsigned int* amax_out = nullptr);
int tc_tiles_per_launch(int B, int Y, int X);
extern int g_fuse_small;
extern int g_fuse_solver_io;
extern int g_wgrad_overlap;
extern int g_wgrad_window_us;
extern int g_wgrad_bg_ctas;          // CTAs of the background weight-gradient launches beside the adjoint conv chain (0 = off)
extern int g_wgrad_bg_chunk;
extern int g_wgrad_issuers;          // MMA-issuing warps of the 3xFP16 weight-gradient kernel (1 or 2)         // unrolled steps per background launch
// 32->32 layers: tensor-core path when enabled (wprep = pre-split weights or NULL for internal scratch), else SIMT
int launch_conv5x5_c32_auto(cudaStream_t st, int B, int Y, int X, const float* in, const float* w, const float* wprep,
                            const float* bias, const float* addend, const float* ref, int act, float slope, float* out,
                            unsigned int* amax_out = nullptr);

}  // namespace sol
