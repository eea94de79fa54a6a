// Host-side engine of libsol_b200.so: plan management, the C ABI (include/sol_b200.h) and the
// unrolled training iteration (reference karman-2d/karman_train.py:393-457 executed by one
// sess.run at :502).  The msteps-unrolled forward sweep, the hand-written adjoint sweep and the
// weight-gradient reduction are enqueued from C++ on one stream — ~50 kernels per unrolled step —
// and (optionally) captured once into a CUDA graph that is replayed every iteration, so the host
// cost per training iteration is one cudaGraphLaunch instead of TensorFlow's graph executor.
#include <math.h>

#include <algorithm>
#include <new>

#include <nvtx3/nvToolsExt.h>

#include "sol_internal.cuh"

SOL_TRACE_TU()

using namespace sol;

namespace sol {
int g_nvtx = 0;      // option "nvtx": NVTX ranges around the stages of the unrolled sweeps (host side; eager launches or graph capture)
}
namespace {
// RAII range: "fwd step 3 / cnn" etc. show up in nsys / `ncu --nvtx`; a no-op unless the option is on
struct NvtxRange {
    bool on;
    explicit NvtxRange(const char* fmt, int a = 0, int b = 0) : on(sol::g_nvtx != 0) {
        if (on) { char buf[96]; snprintf(buf, sizeof(buf), fmt, a, b); nvtxRangePushA(buf); }
    }
    ~NvtxRange() { if (on) nvtxRangePop(); }
};
}  // namespace

namespace {

struct LayerDesc {
    int cin, cout;
    size_t w_off, b_off;   // offsets into the flat Keras-ordered parameter buffer
};

int build_layers(int model, int cin0, std::vector<LayerDesc>& L) {
    L.clear();
    std::vector<std::pair<int, int>> shp;
    if (model == SOL_MODEL_MARS_MOON) {
        shp.push_back({cin0, 32});
        for (int k = 0; k < 10; ++k) shp.push_back({32, 32});
        shp.push_back({32, 2});
    } else if (model == SOL_MODEL_MERCURY) {
        shp.push_back({cin0, 32});
        shp.push_back({32, 64});
        shp.push_back({64, 2});
    } else {
        return SOL_ERR_INVALID;
    }
    size_t off = 0;
    for (auto& s : shp) {
        LayerDesc d;
        d.cin = s.first; d.cout = s.second;
        d.w_off = off; off += (size_t)25 * d.cin * d.cout;
        d.b_off = off; off += d.cout;
        L.push_back(d);
    }
    return SOL_OK;
}

size_t param_count(const std::vector<LayerDesc>& L) { return L.back().b_off + L.back().cout; }

template <typename T>
int upload(T** dst, const T* src, size_t n) {
    SOL_CUDA(cudaMalloc((void**)dst, n * sizeof(T)));
    SOL_CUDA(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return SOL_OK;
}

// Build the multigrid hierarchy of the pressure operator on the host (see sol_cg_mg.cu).
int build_mg(sol_plan* p, const std::vector<unsigned char>& act0) {
    sol_mg& h = p->mg;
    h.valid = false;
    std::vector<std::vector<unsigned char>> act;
    std::vector<int> LY, LX;
    act.push_back(act0); LY.push_back(p->Y); LX.push_back(p->X);
    while (LX.back() > 4 && LX.back() % 2 == 0 && LY.back() % 2 == 0 && (int)act.size() < MG_MAX_LEVELS) {
        const int Yf = LY.back(), Xf = LX.back(), Yc = Yf / 2, Xc = Xf / 2;
        const std::vector<unsigned char>& af = act.back();
        std::vector<unsigned char> ac((size_t)Yc * Xc);
        for (int j = 0; j < Yc; ++j)
            for (int i = 0; i < Xc; ++i) {
                const int s = af[(2 * j) * Xf + 2 * i] + af[(2 * j) * Xf + 2 * i + 1] + af[(2 * j + 1) * Xf + 2 * i] + af[(2 * j + 1) * Xf + 2 * i + 1];
                ac[(size_t)j * Xc + i] = s >= 2 ? 1 : 0;
            }
        act.push_back(ac); LY.push_back(Yc); LX.push_back(Xc);
    }
    const int nlev = (int)act.size();
    if (nlev < 2 || LY.back() * LX.back() > 64) return SOL_OK;   // not supported: plain CG is used
    auto diag_of = [&](int l, int j, int i) -> float {
        auto acc = [&](int jj, int ii) -> float {
            if (jj < 0 || jj >= LY[l] || ii < 0 || ii >= LX[l]) return 1.0f;
            return act[l][(size_t)jj * LX[l] + ii] ? 1.0f : 0.0f;
        };
        const float d = acc(j - 1, i) + acc(j + 1, i) + acc(j, i - 1) + acc(j, i + 1);
        return d < 1.0f ? 1.0f : d;
    };
    std::vector<float> dinv, diag;
    for (int l = 1; l < nlev - 1; ++l) {
        h.coff[l] = (int)dinv.size();
        for (int j = 0; j < LY[l]; ++j)
            for (int i = 0; i < LX[l]; ++i) {
                const float d = diag_of(l, j, i);
                diag.push_back(d);
                dinv.push_back(act[l][(size_t)j * LX[l] + i] ? -h.omega / d : 0.0f);
            }
    }
    // dense inverse on the coarsest level (active cells only), Gauss-Jordan with partial pivoting in double
    const int lc = nlev - 1, Yc = LY[lc], Xc = LX[lc], Nc = Yc * Xc;
    std::vector<double> A((size_t)Nc * Nc, 0.0), Inv((size_t)Nc * Nc, 0.0);
    for (int j = 0; j < Yc; ++j)
        for (int i = 0; i < Xc; ++i) {
            const int c = j * Xc + i;
            Inv[(size_t)c * Nc + c] = 1.0;
            if (!act[lc][c]) { A[(size_t)c * Nc + c] = 1.0; continue; }
            A[(size_t)c * Nc + c] = -(double)diag_of(lc, j, i);
            const int nb[4][2] = {{j - 1, i}, {j + 1, i}, {j, i - 1}, {j, i + 1}};
            for (auto& q : nb)
                if (q[0] >= 0 && q[0] < Yc && q[1] >= 0 && q[1] < Xc && act[lc][q[0] * Xc + q[1]]) A[(size_t)c * Nc + q[0] * Xc + q[1]] = 1.0;
        }
    for (int col = 0; col < Nc; ++col) {
        int piv = col;
        for (int r = col + 1; r < Nc; ++r)
            if (fabs(A[(size_t)r * Nc + col]) > fabs(A[(size_t)piv * Nc + col])) piv = r;
        if (fabs(A[(size_t)piv * Nc + col]) < 1e-12) return SOL_OK;   // singular: leave MG disabled
        if (piv != col)
            for (int k = 0; k < Nc; ++k) { std::swap(A[(size_t)piv * Nc + k], A[(size_t)col * Nc + k]); std::swap(Inv[(size_t)piv * Nc + k], Inv[(size_t)col * Nc + k]); }
        const double d = A[(size_t)col * Nc + col];
        for (int k = 0; k < Nc; ++k) { A[(size_t)col * Nc + k] /= d; Inv[(size_t)col * Nc + k] /= d; }
        for (int r = 0; r < Nc; ++r) {
            if (r == col) continue;
            const double f = A[(size_t)r * Nc + col];
            if (f == 0.0) continue;
            for (int k = 0; k < Nc; ++k) { A[(size_t)r * Nc + k] -= f * A[(size_t)col * Nc + k]; Inv[(size_t)r * Nc + k] -= f * Inv[(size_t)col * Nc + k]; }
        }
    }
    std::vector<float> cinv((size_t)Nc * Nc);
    for (int r = 0; r < Nc; ++r)
        for (int k = 0; k < Nc; ++k) cinv[(size_t)r * Nc + k] = (act[lc][r] && act[lc][k]) ? (float)Inv[(size_t)r * Nc + k] : 0.0f;
    if (dinv.empty()) { dinv.push_back(0.f); diag.push_back(1.f); }
    SOL_TRY(upload(&h.dinv, dinv.data(), dinv.size()));
    SOL_TRY(upload(&h.diag, diag.data(), diag.size()));
    SOL_TRY(upload(&h.cinv, cinv.data(), cinv.size()));
    h.nlev = nlev;
    for (int l = 0; l < nlev; ++l) { h.LY[l] = LY[l]; h.LX[l] = LX[l]; }
    h.valid = true;
    return SOL_OK;
}

}  // namespace

// =================================================================================================
// misc
// =================================================================================================
extern "C" int sol_abi_version(void) { return SOL_ABI_VERSION; }
extern "C" const char* sol_last_error_string(void) { return sol::g_err; }
extern "C" unsigned long long sol_launch_count(void) { return sol::g_launches.load(); }

// ---- chain trace (diagnostics, see sol_internal.cuh) ----
namespace sol {
bool g_trace_names = false;
static std::vector<void (*)(const TraceCtl&)>& trace_setters() { static std::vector<void (*)(const TraceCtl&)> v; return v; }
static std::vector<const void*>& trace_launches() { static std::vector<const void*> v; return v; }
void trace_register(void (*setter)(const TraceCtl&)) { trace_setters().push_back(setter); }
void trace_record_launch(const void* kern) { trace_launches().push_back(kern); }
}  // namespace sol

// buf: device array of `cap` uint64 time stamps, count: device counter (zeroed by the caller); null buf switches tracing off.
extern "C" void sol_debug_chain_trace(unsigned long long* buf, unsigned int* count, unsigned int cap) {
    sol::TraceCtl c{buf, count, cap};
    for (auto f : sol::trace_setters()) f(c);
    cudaDeviceSynchronize();
}
// record_names = 1: remember the kernel of every launch from now on (clears the list); the list is read back as
// newline-separated names in launch order (returns the number of launches recorded).
extern "C" int sol_debug_chain_names(int record_names, char* out, int cap) {
    if (record_names >= 0) { sol::g_trace_names = record_names != 0; if (record_names) sol::trace_launches().clear(); }
    int pos = 0;
    if (out && cap > 0) {
        for (const void* k : sol::trace_launches()) {
            const char* name = nullptr;
            if (cudaFuncGetName(&name, k) != cudaSuccess || !name) name = "?";
            const int n = (int)strlen(name);
            if (pos + n + 2 > cap) break;
            memcpy(out + pos, name, n); pos += n; out[pos++] = '\n';
        }
        out[pos] = 0;
    }
    return (int)sol::trace_launches().size();
}

// Every setter that shapes the launch sequence of an unrolled iteration (process-wide options, per-plan solver settings, the
// Burgers scalars) bumps this counter; a captured graph is replayed only while the counter still has the value it was
// captured at, otherwise the iteration is run eagerly and captured again.
static std::atomic<unsigned long long> g_cfg_epoch{1};
static inline void cfg_changed() { g_cfg_epoch.fetch_add(1, std::memory_order_relaxed); }

extern "C" size_t sol_model_param_count(int model, int cin0) {
    std::vector<LayerDesc> L;
    if (build_layers(model, cin0, L) != SOL_OK) return 0;
    return param_count(L);
}

// =================================================================================================
// plan
// =================================================================================================
extern "C" int sol_plan_create(int Y, int X, int B_max, float dx, int boundary, const unsigned char* solid, const float* inflow,
                               const float* bc_mask_y, const float* bc_val_y, sol_plan** out) {
    SOL_CHECK(out != nullptr, "sol_plan_create: out is NULL");
    SOL_CHECK(Y >= 4 && X >= 4 && B_max >= 1 && dx > 0.f, "sol_plan_create: bad geometry");
    SOL_CHECK(boundary == SOL_BOUNDARY_OPEN || boundary == SOL_BOUNDARY_PERIODIC, "sol_plan_create: bad boundary");
    SOL_CHECK((bc_mask_y == nullptr) == (bc_val_y == nullptr), "sol_plan_create: bc_mask_y and bc_val_y go together");
    if (boundary == SOL_BOUNDARY_PERIODIC)
        SOL_CHECK(!solid && !inflow && !bc_mask_y, "sol_plan_create: periodic plans take no masks");
    int dev = 0;
    SOL_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    SOL_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10) return fail(SOL_ERR_UNSUPPORTED, "libsol_b200 requires a Blackwell (sm_100a) device");
    sol_plan* p = new (std::nothrow) sol_plan();
    if (!p) return fail(SOL_ERR_INVALID, "out of host memory");
    p->Y = Y; p->X = X; p->B_max = B_max; p->dx = dx; p->boundary = boundary; p->sm_count = prop.multiProcessorCount;
    if (boundary == SOL_BOUNDARY_OPEN) {
        const size_t NC = p->NC(), NY = p->NY(), NX = p->NX();
        std::vector<unsigned char> act(NC);
        std::vector<float> diag(NC), my(NY), mx(NX);
        auto acc = [&](int j, int i) -> float {
            if (j < 0 || j >= Y || i < 0 || i >= X) return 1.0f;   // outside an OPEN domain counts as accessible
            return (solid && solid[(size_t)j * X + i]) ? 0.0f : 1.0f;
        };
        for (int j = 0; j < Y; ++j)
            for (int i = 0; i < X; ++i) {
                act[(size_t)j * X + i] = acc(j, i) > 0.f ? 1 : 0;
                const float d = acc(j - 1, i) + acc(j + 1, i) + acc(j, i - 1) + acc(j, i + 1);
                diag[(size_t)j * X + i] = d < 1.0f ? 1.0f : d;
            }
        for (int j = 0; j <= Y; ++j)
            for (int i = 0; i < X; ++i) my[(size_t)j * X + i] = fminf(acc(j - 1, i), acc(j, i));
        for (int j = 0; j < Y; ++j)
            for (int i = 0; i <= X; ++i) mx[(size_t)j * (X + 1) + i] = fminf(acc(j, i - 1), acc(j, i));
        int rc = upload(&p->active, act.data(), NC);
        if (rc == SOL_OK) rc = upload(&p->diag, diag.data(), NC);
        if (rc == SOL_OK) rc = upload(&p->face_my, my.data(), NY);
        if (rc == SOL_OK) rc = upload(&p->face_mx, mx.data(), NX);
        if (rc == SOL_OK && inflow) rc = upload(&p->inflow, inflow, NC);
        if (rc == SOL_OK && bc_mask_y) rc = upload(&p->bc_mask_y, bc_mask_y, NY);
        if (rc == SOL_OK && bc_val_y) rc = upload(&p->bc_val_y, bc_val_y, NY);
        if (rc == SOL_OK) rc = build_mg(p, act);
        p->h_active = act; p->h_diag = diag;
        if (rc != SOL_OK) { sol_plan_destroy(p); return rc; }
    }
    *out = p;
    return SOL_OK;
}

extern "C" int sol_plan_destroy(sol_plan* p) {
    if (!p) return SOL_OK;
    cudaFree(p->active); cudaFree(p->diag); cudaFree(p->face_my); cudaFree(p->face_mx);
    cudaFree(p->inflow); cudaFree(p->bc_mask_y); cudaFree(p->bc_val_y);
    cudaFree(p->mg.dinv); cudaFree(p->mg.diag); cudaFree(p->mg.cinv);
    cudaFree(p->cg_any_scratch);
    direct_free(p);
    delete p;
    return SOL_OK;
}

extern "C" int sol_plan_set_cg(sol_plan* p, float tol_abs, float tol_rel, int max_it, int cluster) {
    SOL_CHECK(p != nullptr, "sol_plan_set_cg: plan is NULL");
    SOL_CHECK(tol_abs >= 0.f && tol_rel >= 0.f && max_it >= 0, "sol_plan_set_cg: negative tolerance / iteration cap");
    SOL_CHECK(cluster == 0 || cluster == 1 || cluster == 2 || cluster == 4 || cluster == 8, "sol_plan_set_cg: cluster must be 0,1,2,4,8");
    cfg_changed();
    p->tol_abs = tol_abs; p->tol_rel = tol_rel; p->max_it = max_it; p->cluster = cluster;
    return SOL_OK;
}

extern "C" int sol_plan_set_option(sol_plan* p, const char* name, int value) {
    SOL_CHECK(p != nullptr && name != nullptr, "sol_plan_set_option: NULL pointer");
    cfg_changed();
    if (strcmp(name, "cg_rows") == 0) {
        SOL_CHECK(value == 0 || value == 2 || value == 4 || value == 8 || value == 16, "cg_rows must be 0,2,4,8,16");
        p->cg_rows = value;
        return SOL_OK;
    }
    if (strcmp(name, "mg_variant") == 0) {
        SOL_CHECK(value == 0 || value == 2, "mg_variant must be 0 or 2");
        p->mg_variant = value;
        return SOL_OK;
    }
    if (strcmp(name, "cg_precond") == 0) {
        SOL_CHECK(value == 0 || value == 1, "cg_precond must be 0 or 1");
        p->cg_precond = value;
        return SOL_OK;
    }
    if (strcmp(name, "direct_solve") == 0) {
        SOL_CHECK(value == 0 || value == 1, "direct_solve must be 0 or 1");
        p->direct_solve = value;
        if (value) (void)direct_active(p);      // precompute now (outside any stream capture)
        return SOL_OK;
    }
    return fail(SOL_ERR_INVALID, "sol_plan_set_option: unknown option");
}

extern "C" int sol_plan_query(sol_plan* p, const char* name, int* value) {
    SOL_CHECK(p != nullptr && name != nullptr && value != nullptr, "sol_plan_query: NULL pointer");
    if (strcmp(name, "direct_active") == 0) { *value = direct_for_batch(p, 1) ? 1 : 0; return SOL_OK; }
    if (strcmp(name, "direct_rows") == 0) { *value = direct_active(p) ? p->dir.kp : 0; return SOL_OK; }
    if (strcmp(name, "sm_count") == 0) { *value = p->sm_count; return SOL_OK; }
    return fail(SOL_ERR_INVALID, "sol_plan_query: unknown name");
}

extern "C" int sol_set_option(const char* name, int value) {
    SOL_CHECK(name != nullptr, "sol_set_option: NULL name");
    cfg_changed();
    if (strcmp(name, "conv_path") == 0) {
        SOL_CHECK(value >= 0 && value <= 3, "conv_path must be 0,1,2,3");
        sol::g_conv_path = value == 0 ? 2 : value;
        return SOL_OK;
    }
    if (strcmp(name, "wgrad_path") == 0) {
        SOL_CHECK(value >= 0 && value <= 3, "wgrad_path must be 0,1,2,3");
        sol::g_wgrad_path = value == 0 ? 2 : value;
        return SOL_OK;
    }
    if (strcmp(name, "fuse_solver_io") == 0) {
        sol::g_fuse_solver_io = value ? 1 : 0;
        return SOL_OK;
    }
    if (strcmp(name, "fuse_small") == 0) {
        sol::g_fuse_small = value ? 1 : 0;
        return SOL_OK;
    }
    if (strcmp(name, "wgrad_window_us") == 0) {
        SOL_CHECK(value >= 0 && value <= 10000, "wgrad_window_us out of range");
        sol::g_wgrad_window_us = value;
        return SOL_OK;
    }
    if (strcmp(name, "wgrad_bg_ctas") == 0) {
        SOL_CHECK(value >= 0 && value <= 148, "wgrad_bg_ctas out of range");
        sol::g_wgrad_bg_ctas = value;
        return SOL_OK;
    }
    if (strcmp(name, "wgrad_bg_chunk") == 0) {
        SOL_CHECK(value >= 1 && value <= 64, "wgrad_bg_chunk out of range");
        sol::g_wgrad_bg_chunk = value;
        return SOL_OK;
    }
    if (strcmp(name, "wgrad_issuers") == 0) {
        SOL_CHECK(value == 1 || value == 2, "wgrad_issuers must be 1 or 2");
        sol::g_wgrad_issuers = value;
        return SOL_OK;
    }
    if (strcmp(name, "wgrad_overlap") == 0) {
        sol::g_wgrad_overlap = value ? 1 : 0;
        return SOL_OK;
    }
    if (strcmp(name, "conv_variant") == 0) {
        SOL_CHECK(value >= 0 && value <= 2, "conv_variant must be 0,1,2");
        sol::g_conv_variant = value;
        return SOL_OK;
    }
    if (strcmp(name, "pdl") == 0) {
        sol::g_pdl = value ? 1 : 0;
        return SOL_OK;
    }
    if (strcmp(name, "deterministic") == 0) {
        sol::g_deterministic = value ? 1 : 0;
        return SOL_OK;
    }
    if (strcmp(name, "thin_path") == 0) {
        SOL_CHECK(value == 0 || value == 1, "thin_path must be 0 (row-pair kernels) or 1 (first-generation kernels)");
        sol::g_thin_path = value;
        return SOL_OK;
    }
    if (strcmp(name, "thin_cap") == 0) {
        SOL_CHECK(value >= 0 && value <= 7, "thin_cap is a 3-bit mask");
        sol::set_thin_cap_mask(value);
        return SOL_OK;
    }
    if (strcmp(name, "fuse_stencil") == 0) {
        sol::g_fuse_stencil = value ? 1 : 0;
        return SOL_OK;
    }
    if (strcmp(name, "nvtx") == 0) {
        sol::g_nvtx = value ? 1 : 0;
        return SOL_OK;
    }
    if (strcmp(name, "tc_base_offset_mode") == 0) {
        sol::g_tc_base_offset_mode = value ? 1 : 0;
        return SOL_OK;
    }
    return fail(SOL_ERR_INVALID, "sol_set_option: unknown option");
}

#define SOL_PLAN_B(p, B)                                               \
    SOL_CHECK((p) != nullptr, "plan is NULL");                         \
    SOL_CHECK((B) >= 1 && (B) <= (p)->B_max, "batch size out of range for this plan")

// =================================================================================================
// stage entry points
// =================================================================================================
extern "C" int sol_diffuse_bc(sol_plan* p, void* stream, int B, const float* re, float dt, float res, const float* vy, const float* vx,
                              float* vy_out, float* vx_out) {
    SOL_PLAN_B(p, B);
    SOL_CHECK(re && vy && vx && vy_out && vx_out, "sol_diffuse_bc: NULL pointer");
    return launch_diffuse_bc(p, (cudaStream_t)stream, B, re, dt, res, vy, vx, vy_out, vx_out);
}

extern "C" int sol_diffuse_bc_bwd(sol_plan* p, void* stream, int B, const float* re, float dt, float res, const float* gy, const float* gx,
                                  float* gy_in, float* gx_in) {
    SOL_PLAN_B(p, B);
    SOL_CHECK(re && gy && gx && gy_in && gx_in, "sol_diffuse_bc_bwd: NULL pointer");
    return launch_diffuse_bc_bwd(p, (cudaStream_t)stream, B, re, dt, res, gy, gx, gy_in, gx_in, nullptr, nullptr);
}

extern "C" int sol_diffuse_advect(sol_plan* p, void* stream, int B, const float* re, float dt, float res, const float* vy, const float* vx,
                                  const float* rho, float* vy1, float* vx1, float* vy2, float* vx2, float* rho_out) {
    SOL_PLAN_B(p, B);
    SOL_CHECK(re && vy && vx && vy1 && vx1 && vy2 && vx2, "sol_diffuse_advect: NULL pointer");
    SOL_CHECK((rho == nullptr) == (rho_out == nullptr), "sol_diffuse_advect: rho and rho_out go together");
    return launch_diffuse_advect(p, (cudaStream_t)stream, B, re, dt, res, vy, vx, rho, vy1, vx1, vy2, vx2, rho_out);
}

extern "C" int sol_advect(sol_plan* p, void* stream, int B, float dt, const float* vy, const float* vx, const float* rho, float* vy_out,
                          float* vx_out, float* rho_out) {
    SOL_PLAN_B(p, B);
    SOL_CHECK(vy && vx && vy_out && vx_out, "sol_advect: NULL pointer");
    SOL_CHECK((rho == nullptr) == (rho_out == nullptr), "sol_advect: rho and rho_out go together");
    return launch_advect(p, (cudaStream_t)stream, B, dt, vy, vx, rho, vy_out, vx_out, rho_out);
}

extern "C" int sol_advect_bwd(sol_plan* p, void* stream, int B, float dt, const float* vy, const float* vx, const float* gy_out,
                              const float* gx_out, float* gy, float* gx) {
    SOL_PLAN_B(p, B);
    SOL_CHECK(vy && vx && gy_out && gx_out && gy && gx, "sol_advect_bwd: NULL pointer");
    return launch_advect_bwd(p, (cudaStream_t)stream, B, dt, vy, vx, gy_out, gx_out, gy, gx);
}

extern "C" int sol_pressure_solve(sol_plan* p, void* stream, int B, const float* div, float* pr, int* iters) {
    SOL_PLAN_B(p, B);
    SOL_CHECK(div && pr, "sol_pressure_solve: NULL pointer");
    return launch_cg(p, (cudaStream_t)stream, B, 0, div, pr, nullptr, nullptr, nullptr, nullptr, iters);
}

extern "C" int sol_project(sol_plan* p, void* stream, int B, const float* vy, const float* vx, float* vy_out, float* vx_out, float* p_out,
                           int* iters) {
    SOL_PLAN_B(p, B);
    SOL_CHECK(vy && vx && vy_out && vx_out, "sol_project: NULL pointer");
    return launch_cg(p, (cudaStream_t)stream, B, 1, nullptr, p_out, vy, vx, vy_out, vx_out, iters);
}

extern "C" int sol_divergence(sol_plan* p, void* stream, int B, const float* vy, const float* vx, float* div) {
    SOL_PLAN_B(p, B);
    SOL_CHECK(vy && vx && div, "sol_divergence: NULL pointer");
    SOL_CHECK(p->boundary == SOL_BOUNDARY_OPEN, "sol_divergence: OPEN plans only");
    return launch_divergence(p, (cudaStream_t)stream, B, vy, vx, div);
}

extern "C" int sol_step_fwd(sol_plan* p, void* stream, int B, const float* re, float dt, float res, const float* rho_in, const float* vy_in,
                            const float* vx_in, float* rho_out, float* vy_out, float* vx_out, float* p_out, float* vy1, float* vx1,
                            float* scratch_vy, float* scratch_vx, int* iters) {
    SOL_PLAN_B(p, B);
    SOL_CHECK(re && vy_in && vx_in && vy_out && vx_out && scratch_vy && scratch_vx, "sol_step_fwd: NULL pointer");
    SOL_CHECK((rho_in == nullptr) == (rho_out == nullptr), "sol_step_fwd: rho_in and rho_out go together");
    SOL_CHECK((vy1 == nullptr) == (vx1 == nullptr), "sol_step_fwd: vy1 and vx1 go together");
    cudaStream_t st = (cudaStream_t)stream;
    float* d_vy = vy1 ? vy1 : vy_out;
    float* d_vx = vx1 ? vx1 : vx_out;
    if (sol::g_fuse_stencil && p->boundary == SOL_BOUNDARY_OPEN && d_vy != vy_in && d_vx != vx_in) {
        SOL_TRY(launch_diffuse_advect(p, st, B, re, dt, res, vy_in, vx_in, rho_in, d_vy, d_vx, scratch_vy, scratch_vx, rho_out));
    } else {
        SOL_TRY(launch_diffuse_bc(p, st, B, re, dt, res, vy_in, vx_in, d_vy, d_vx));
        SOL_TRY(launch_advect(p, st, B, dt, d_vy, d_vx, rho_in, scratch_vy, scratch_vx, rho_out));
    }
    return launch_cg(p, st, B, 1, nullptr, p_out, scratch_vy, scratch_vx, vy_out, vx_out, iters);
}

extern "C" int sol_step_bwd(sol_plan* p, void* stream, int B, const float* re, float dt, float res, const float* vy1, const float* vx1,
                            const float* gy_out, const float* gx_out, float* gy_in, float* gx_in, float* scratch_vy, float* scratch_vx,
                            int* iters) {
    SOL_PLAN_B(p, B);
    SOL_CHECK(re && vy1 && vx1 && gy_out && gx_out && gy_in && gx_in && scratch_vy && scratch_vx, "sol_step_bwd: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    // the projection is self-adjoint
    SOL_TRY(launch_cg(p, st, B, 1, nullptr, nullptr, gy_out, gx_out, gy_in, gx_in, iters));
    SOL_TRY(launch_advect_bwd(p, st, B, dt, vy1, vx1, gy_in, gx_in, scratch_vy, scratch_vx));
    return launch_diffuse_bc_bwd(p, st, B, re, dt, res, scratch_vy, scratch_vx, gy_in, gx_in, nullptr, nullptr);
}

extern "C" int sol_burgers_step(sol_plan* p, void* stream, int B, float dt, float viscosity, const float* ky, const float* kx,
                                const float* vy, const float* vx, const float* fy, const float* fx, float* vy_out, float* vx_out,
                                float* scratch_vy, float* scratch_vx) {
    SOL_PLAN_B(p, B);
    SOL_CHECK(p->boundary == SOL_BOUNDARY_PERIODIC, "sol_burgers_step: PERIODIC plans only");
    SOL_CHECK(vy && vx && vy_out && vx_out && scratch_vy && scratch_vx, "sol_burgers_step: NULL pointer");
    SOL_CHECK((ky == nullptr) == (kx == nullptr) && (fy == nullptr) == (fx == nullptr), "sol_burgers_step: kernels / forces come in pairs");
    cudaStream_t st = (cudaStream_t)stream;
    SOL_TRY(launch_advect(p, st, B, dt, vy, vx, nullptr, scratch_vy, scratch_vx, nullptr));
    return launch_burgers_diffuse(p, st, B, viscosity * dt, ky, kx, scratch_vy, scratch_vx, fy, fx, dt, vy_out, vx_out);
}

extern "C" int sol_burgers_step_bwd(sol_plan* p, void* stream, int B, float dt, float viscosity, const float* ky, const float* kx,
                                    const float* vy, const float* vx, const float* gy_out, const float* gx_out, float* gy, float* gx,
                                    float* scratch_vy, float* scratch_vx) {
    SOL_PLAN_B(p, B);
    SOL_CHECK(p->boundary == SOL_BOUNDARY_PERIODIC, "sol_burgers_step_bwd: PERIODIC plans only");
    SOL_CHECK(vy && vx && gy_out && gx_out && gy && gx && scratch_vy && scratch_vx, "sol_burgers_step_bwd: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    // the periodic diffusion operator is symmetric (real, even kernel): its adjoint is itself
    SOL_TRY(launch_burgers_diffuse(p, st, B, viscosity * dt, ky, kx, gy_out, gx_out, nullptr, nullptr, 0.f, scratch_vy, scratch_vx));
    return launch_advect_bwd(p, st, B, dt, vy, vx, scratch_vy, scratch_vx, gy, gx);
}

extern "C" int sol_conv5x5(void* stream, int B, int Y, int X, int Cin, int Cout, const float* in, const float* w, const float* bias,
                           const float* addend, const float* ref, int act, float slope, float* out) {
    SOL_CHECK(in && w && out && B >= 1 && Y >= 1 && X >= 1, "sol_conv5x5: bad arguments");
    if (Cin == 32 && Cout == 32)
        return launch_conv5x5_c32_auto((cudaStream_t)stream, B, Y, X, in, w, nullptr, bias, addend, ref, act, slope, out);
    return launch_conv5x5((cudaStream_t)stream, B, Y, X, Cin, Cout, in, w, bias, addend, ref, act, slope, out);
}

extern "C" size_t sol_conv5x5_split_floats(void) { return tc_weights_floats(); }

extern "C" int sol_conv5x5_split_weights(void* stream, const float* w, float* wsplit) {
    SOL_CHECK(w && wsplit, "sol_conv5x5_split_weights: NULL pointer");
    return launch_split_weights((cudaStream_t)stream, w, wsplit);
}

extern "C" int sol_conv5x5_c32_presplit(void* stream, int B, int Y, int X, const float* in, const float* wsplit, const float* bias,
                                        const float* addend, const float* ref, int act, float slope, float* out, int weights_settled) {
    SOL_CHECK(in && wsplit && out && B >= 1 && Y >= 1 && X >= 1, "sol_conv5x5_c32_presplit: bad arguments");
    SOL_CHECK(!(act == SOL_ACT_DLRELU && !ref), "sol_conv5x5_c32_presplit: SOL_ACT_DLRELU needs ref");
    SOL_CHECK(conv_path_is_tc(), "sol_conv5x5_c32_presplit: option conv_path selects the SIMT kernels");
    return launch_conv5x5_c32_presplit((cudaStream_t)stream, B, Y, X, in, wsplit, bias, addend, ref, act, slope, out, weights_settled != 0);
}

extern "C" int sol_conv5x5_flip_weights(void* stream, int Cin, int Cout, const float* w, float* wT) {
    SOL_CHECK(w && wT && Cin >= 1 && Cout >= 1, "sol_conv5x5_flip_weights: bad arguments");
    return launch_flip_weights((cudaStream_t)stream, Cin, Cout, w, wT);
}

extern "C" size_t sol_conv5x5_wgrad_workspace(int Cin, int Cout) { return wgrad_workspace_floats(Cin, Cout); }

extern "C" int sol_conv5x5_wgrad(void* stream, int B, int Y, int X, int Cin, int Cout, const float* in, const float* g_out, float* dW,
                                 float* db, int accumulate, float* partials) {
    SOL_CHECK(in && g_out && dW && db, "sol_conv5x5_wgrad: NULL pointer");
    if (Cin == 32 && Cout == 32 && sol::g_wgrad_path >= 2 && Y % 16 == 0 && X % 8 == 0) {
        // tensor-core GEMM over the pixels of this one batch (the engine defers it over all unrolled steps)
        SOL_CHECK(partials != nullptr, "sol_conv5x5_wgrad: partials workspace required");
        cudaStream_t st = (cudaStream_t)stream;
        int dev = 0, sms = 148, nctas = 0;
        SOL_CUDA(cudaGetDevice(&dev));
        SOL_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const size_t stride = (size_t)B * Y * X * 32;
        if (sol::g_wgrad_path == 3) {
            SOL_TRY(launch_wgrad_c32_tc(st, sms, 1, B, Y, X, in, stride, g_out, stride, partials, &nctas));
        } else {
            // stand-alone call: nobody tracked the operand maxima, measure them (stream-ordered scratch slots)
            static unsigned int* slots = nullptr;
            if (!slots) SOL_CUDA(cudaMalloc((void**)&slots, 2 * sizeof(unsigned int)));
            SOL_CUDA(cudaMemsetAsync(slots, 0, 2 * sizeof(unsigned int), st));
            SOL_TRY(launch_amax(st, in, stride, slots));
            SOL_TRY(launch_amax(st, g_out, stride, slots + 1));
            SOL_TRY(launch_wgrad_c32_h(st, sms, 1, B, Y, X, in, stride, g_out, stride, slots, slots + 1, partials, &nctas));
        }
        return launch_wgrad_finalize_n(st, nctas, partials, dW, db, accumulate);
    }
    return launch_wgrad((cudaStream_t)stream, B, Y, X, Cin, Cout, in, g_out, dW, db, accumulate, partials, true);
}

extern "C" int sol_to_feature(sol_plan* p, void* stream, int B, const float* vy, const float* vx, const float* re, float sy, float sx,
                              float sr, float* feat) {
    SOL_PLAN_B(p, B);
    SOL_CHECK(vy && vx && re && feat, "sol_to_feature: NULL pointer");
    SOL_CHECK(sy > 0.f && sx > 0.f && sr > 0.f, "sol_to_feature: sigmas must be positive");
    return launch_to_feature(p, (cudaStream_t)stream, B, vy, vx, re, sy, sx, sr, feat);
}

extern "C" int sol_adam_tf1(void* stream, size_t n, float* theta, const float* grad, float* m, float* v, int t, float lr, float beta1,
                            float beta2, float eps, float grad_scale) {
    SOL_CHECK(theta && grad && m && v && t >= 1, "sol_adam_tf1: bad arguments");
    const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, t)) / (1.0 - pow((double)beta1, t));
    return launch_adam((cudaStream_t)stream, n, theta, grad, m, v, (float)lr_t, beta1, beta2, eps, grad_scale);
}

// =================================================================================================
// the unrolled training iteration
// =================================================================================================
struct StepStash {
    float *vy1, *vx1;     // post-BC velocity (input of the advection) — advection adjoint
    float* feat;          // CNN input features
    float* acts[11];      // a0, t1, a1, ..., t5, a5
    float *gl_vy, *gl_vx; // d(loss)/d(corrected state) of this step
};

struct sol_unroll {
    sol_plan* plan = nullptr;
    sol_unroll_cfg cfg;
    std::vector<LayerDesc> L;
    size_t nparams = 0;
    std::vector<StepStash> stash;
    // transients
    float *sA_vy, *sA_vx, *sB_vy, *sB_vx, *rhoA, *rhoB;
    float *vy2, *vx2, *vy3, *vx3;
    float* corr;
    float *G_vy[2], *G_vx[2], *H_vy, *H_vx, *H2_vy, *H2_vx, *K_vy, *K_vx;
    float *g_corr, *g_feat, *gbuf[3];
    float* wT;
    float *wprep_fwd, *wprep_bwd;   // [10 layers][2*25*32*32] pre-split tensor-core weights
    float* gst;        // deferred weight gradient: [10 layers][msteps][B,Y,X,32] output-gradient stash
    float* g0_st;      // [msteps][B,Y,X,32] output gradient of layer 0
    bool deferred_wgrad = false;   // decided per backward sweep: option wgrad_path >= 2 and the grid tiles evenly (Y%16, X%8)
    float* thin_part = nullptr;    // deterministic mode: private CTA slots of the thin-layer weight gradients
    float* loss_part = nullptr;    // deterministic mode: [msteps][loss_part_stride] per-CTA loss partials
    int loss_part_stride = 0;
    unsigned int* amax = nullptr;  // [24] running max|x| (bit patterns): [l] input activations of layer l, [12 + l] its output gradients (l = 1..10)
    bool track_amax = false;       // the 3xFP16 conv kernels of this sweep keep them up to date (else the weight-gradient launches measure them)
    float* gcorr_st;   // [msteps][B,Y,X,2]  output gradient of layer 11
    size_t nA = 0;
    float* partials;   // [n_c32][WG_MAX_CTAS][25632]
    size_t partial_stride = 0;
    int* iters;
    float* re_buf;     // private copy of Re[B]: the adjoint sweep must not depend on the caller keeping `re` alive
    bool forward_done = false, have_loss = false;
    const float* last_re = nullptr;
    // Burgers scene (PERIODIC plans, sol_unroll_set_burgers): viscosity, real-space diffusion kernels, per-step forces
    bool burgers_set = false;
    float visc = 0.1f, sig_fy = 1.f, sig_fx = 1.f;
    const float *bk_y = nullptr, *bk_x = nullptr, *bf_vy = nullptr, *bf_vx = nullptr;
    unsigned long long graph_kernels = 0;
    // CUDA graph cache
    cudaGraphExec_t gexec = nullptr;
    const void* gkey[13] = {nullptr};
    unsigned long long gepoch = 0;      // g_cfg_epoch the graph was captured at
    int warm = 0;
    // graph work runs on a private non-blocking stream (the caller's may be the legacy default
    // stream, which cannot be captured); fork/join with events keeps the caller's stream ordering
    cudaStream_t gstream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaStream_t sstream = nullptr;                    // side stream: weight-gradient GEMMs in the shadow of the adjoint solves
    cudaEvent_t ev_wfork = nullptr, ev_wjoin = nullptr;
};

namespace {

struct Carver {
    char* base; size_t off = 0;
    explicit Carver(void* b) : base((char*)b) {}
    template <typename T> T* take(size_t n) {
        off = align_up(off, 256);
        T* p = base ? (T*)(base + off) : nullptr;
        off += n * sizeof(T);
        return p;
    }
};

int carve(sol_unroll* u, void* ws, size_t* total) {
    const sol_plan* p = u->plan;
    const sol_unroll_cfg& c = u->cfg;
    const size_t B = c.B, NC = p->NC() * B, NY = p->NY() * B, NX = p->NX() * B;
    const size_t nA = NC * 32;
    const bool mercury = c.model == SOL_MODEL_MERCURY;
    Carver cv(ws);
    u->stash.resize(c.msteps);
    for (int i = 0; i < c.msteps; ++i) {
        StepStash& s = u->stash[i];
        s.vy1 = cv.take<float>(NY); s.vx1 = cv.take<float>(NX);
        s.feat = cv.take<float>(NC * c.cin0);
        if (mercury) {     // model_mercury: relu(conv 32), relu(conv 64)
            s.acts[0] = cv.take<float>(NC * 32); s.acts[1] = cv.take<float>(NC * 64);
            for (int k = 2; k < 11; ++k) s.acts[k] = nullptr;
        } else {
            for (int k = 0; k < 11; ++k) s.acts[k] = cv.take<float>(nA);
        }
        s.gl_vy = cv.take<float>(NY); s.gl_vx = cv.take<float>(NX);
    }
    u->sA_vy = cv.take<float>(NY); u->sA_vx = cv.take<float>(NX);
    u->sB_vy = cv.take<float>(NY); u->sB_vx = cv.take<float>(NX);
    u->rhoA = cv.take<float>(NC); u->rhoB = cv.take<float>(NC);
    u->vy2 = cv.take<float>(NY); u->vx2 = cv.take<float>(NX);
    u->vy3 = cv.take<float>(NY); u->vx3 = cv.take<float>(NX);
    u->corr = cv.take<float>(NC * 2);
    for (int k = 0; k < 2; ++k) { u->G_vy[k] = cv.take<float>(NY); u->G_vx[k] = cv.take<float>(NX); }
    u->H_vy = cv.take<float>(NY); u->H_vx = cv.take<float>(NX);
    u->H2_vy = cv.take<float>(NY); u->H2_vx = cv.take<float>(NX);
    u->K_vy = cv.take<float>(NY); u->K_vx = cv.take<float>(NX);
    u->g_corr = cv.take<float>(NC * 2);
    u->g_feat = cv.take<float>(NC * c.cin0);
    u->wT = cv.take<float>(u->nparams);
    u->nA = nA;
    if (mercury) {      // per-step SIMT weight gradients, no tensor-core operand buffers, no deferred-gradient stash
        u->gbuf[0] = cv.take<float>(NC * 64); u->gbuf[1] = cv.take<float>(nA); u->gbuf[2] = nullptr;
        u->wprep_fwd = u->wprep_bwd = nullptr;
        u->gst = u->g0_st = u->gcorr_st = u->partials = nullptr; u->partial_stride = 0;
    } else {
        for (int k = 0; k < 3; ++k) u->gbuf[k] = cv.take<float>(nA);
        u->wprep_fwd = cv.take<float>(tc_weights_floats() * 10);
        u->wprep_bwd = cv.take<float>(tc_weights_floats() * 10);
        u->gst = cv.take<float>(nA * 10 * c.msteps);
        u->g0_st = cv.take<float>(nA * c.msteps);
        u->gcorr_st = cv.take<float>(NC * 2 * c.msteps);
        u->partial_stride = wgrad_workspace_floats(32, 32);
        u->partials = cv.take<float>(u->partial_stride * 10);
    }
    u->amax = cv.take<unsigned int>(24);
    u->thin_part = cv.take<float>(wgrad_thin_part_floats());
    u->loss_part_stride = p->sm_count;
    u->loss_part = cv.take<float>((size_t)c.msteps * u->loss_part_stride);
    u->iters = cv.take<int>((size_t)2 * c.msteps * c.B);
    u->re_buf = cv.take<float>(c.B);
    *total = align_up(cv.off, 256);
    return SOL_OK;
}

int check_cfg(const sol_plan* p, const sol_unroll_cfg* c) {
    SOL_CHECK(p && c, "unroll: NULL plan / cfg");
    SOL_CHECK(c->msteps >= 1 && c->B >= 1 && c->B <= p->B_max, "unroll: bad msteps / batch");
    SOL_CHECK(c->sig_vy > 0.f && c->sig_vx > 0.f && c->sig_ext > 0.f, "unroll: sigmas must be positive");
    if (p->boundary == SOL_BOUNDARY_OPEN)
        SOL_CHECK(c->cin0 == 3, "unroll: karman features have 3 channels (vy, vx, Re)");
    else
        SOL_CHECK(c->cin0 == 4 || c->cin0 == 2, "unroll: burgers features have 4 channels (vy, vx, fy, fx) or 2 (--noforce)");
    SOL_CHECK(c->model == SOL_MODEL_MARS_MOON || c->model == SOL_MODEL_MERCURY, "unroll: unknown model");
    return SOL_OK;
}

// A 32->32 layer of the sweep (tensor-core path with the weights split at the start of the sweep, else SIMT)
int layer_conv(sol_unroll* u, cudaStream_t st, const float* in, const float* w, const float* wprep, const float* bias, const float* addend,
               const float* ref, int act, float slope, float* out, unsigned int* amax_out = nullptr) {
    const sol_plan* p = u->plan;
    return launch_conv5x5_c32_auto(st, u->cfg.B, p->Y, p->X, in, w, wprep, bias, addend, ref, act, slope, out, amax_out);
}

// ---- CNN forward / backward over the stash of one step (model_mars_moon, karman_train.py:101-138)
int cnn_forward(sol_unroll* u, cudaStream_t st, const float* w, const StepStash& s, float* corr) {
    const sol_plan* p = u->plan;
    const int B = u->cfg.B, Y = p->Y, X = p->X;
    const float a = 0.3f;   // keras LeakyReLU default
    const std::vector<LayerDesc>& L = u->L;
    if (u->cfg.model == SOL_MODEL_MERCURY) {     // karman_train.py:92-99: conv32+relu, conv64+relu, conv2 (relu = leaky slope 0)
        SOL_TRY(launch_conv5x5(st, B, Y, X, L[0].cin, 32, s.feat, w + L[0].w_off, w + L[0].b_off, nullptr, nullptr, SOL_ACT_LRELU, 0.0f, s.acts[0]));
        SOL_TRY(launch_conv5x5(st, B, Y, X, 32, 64, s.acts[0], w + L[1].w_off, w + L[1].b_off, nullptr, nullptr, SOL_ACT_LRELU, 0.0f, s.acts[1]));
        return launch_conv5x5(st, B, Y, X, 64, 2, s.acts[1], w + L[2].w_off, w + L[2].b_off, nullptr, nullptr, SOL_ACT_NONE, 0.0f, corr);
    }
    const bool tc = conv_path_is_tc();
    unsigned int* am = u->track_amax ? u->amax : nullptr;       // am[l]: max|input of layer l| over the sweep
    // (thin layers: the weights were settled before the sweep started, hence before the preceding kernel)
    SOL_TRY(launch_conv5x5(st, B, Y, X, L[0].cin, 32, s.feat, w + L[0].w_off, w + L[0].b_off, nullptr, nullptr, SOL_ACT_LRELU, a, s.acts[0],
                           am ? am + 1 : nullptr, true));
    for (int k = 1; k <= 5; ++k) {
        const LayerDesc& l1 = L[2 * k - 1];
        const LayerDesc& l2 = L[2 * k];
        float* a_prev = s.acts[2 * k - 2];
        float* t_k = s.acts[2 * k - 1];
        float* a_k = s.acts[2 * k];
        const float* p1 = tc ? u->wprep_fwd + tc_weights_floats() * (2 * k - 2) : nullptr;
        const float* p2 = tc ? u->wprep_fwd + tc_weights_floats() * (2 * k - 1) : nullptr;
        SOL_TRY(layer_conv(u, st, a_prev, w + l1.w_off, p1, w + l1.b_off, nullptr, nullptr, SOL_ACT_LRELU, a, t_k, am ? am + 2 * k : nullptr));
        SOL_TRY(layer_conv(u, st, t_k, w + l2.w_off, p2, w + l2.b_off, a_prev, nullptr, SOL_ACT_LRELU, a, a_k,
                           (am && k < 5) ? am + 2 * k + 1 : nullptr));
    }
    return launch_conv5x5(st, B, Y, X, 32, 2, s.acts[10], w + L[11].w_off, w + L[11].b_off, nullptr, nullptr, SOL_ACT_NONE, a, corr, nullptr, true);
}

int cnn_backward(sol_unroll* u, cudaStream_t st, const float* w, float* gw, const StepStash& s, const float* g_corr, float* g_feat,
                 int first, int step) {
    const sol_plan* p = u->plan;
    const int B = u->cfg.B, Y = p->Y, X = p->X;
    const float a = 0.3f;
    const std::vector<LayerDesc>& L = u->L;
    const float* wT = u->wT;
    if (u->cfg.model == SOL_MODEL_MERCURY) {
        // weight gradients accumulate straight into gw (zeroed at the start of the sweep), step by step
        float* g64 = u->gbuf[0]; float* g32 = u->gbuf[1];
        SOL_TRY(launch_wgrad(st, B, Y, X, 64, 2, s.acts[1], g_corr, gw + L[2].w_off, gw + L[2].b_off, 1, nullptr, false));
        SOL_TRY(launch_conv5x5(st, B, Y, X, 2, 64, g_corr, wT + L[2].w_off, nullptr, nullptr, s.acts[1], SOL_ACT_DLRELU, 0.0f, g64));
        SOL_TRY(launch_wgrad(st, B, Y, X, 32, 64, s.acts[0], g64, gw + L[1].w_off, gw + L[1].b_off, 1, nullptr, false));
        SOL_TRY(launch_conv5x5(st, B, Y, X, 64, 32, g64, wT + L[1].w_off, nullptr, nullptr, s.acts[0], SOL_ACT_DLRELU, 0.0f, g32));
        SOL_TRY(launch_wgrad(st, B, Y, X, L[0].cin, 32, s.feat, g32, gw + L[0].w_off, gw + L[0].b_off, 1, nullptr, false));
        return launch_conv5x5(st, B, Y, X, 32, L[0].cin, g32, wT + L[0].w_off, nullptr, nullptr, nullptr, SOL_ACT_NONE, 0.0f, g_feat);
    }
    const bool tc = conv_path_is_tc();
    const bool deferred = u->deferred_wgrad;        // weight gradients in one GEMM per layer after the sweep
    // output-gradient tensor of layer l (1..10) for this step
    auto gout = [&](int l, float* fallback) -> float* {
        return deferred ? u->gst + ((size_t)(l - 1) * u->cfg.msteps + step) * u->nA : fallback;
    };
    float* gS = gout(10, u->gbuf[0]);
    float* spare[3] = {u->gbuf[0], u->gbuf[1], u->gbuf[2]};
    // output layer (32 -> 2)
    if (!deferred)
        SOL_TRY(launch_wgrad(st, B, Y, X, 32, 2, s.acts[10], g_corr, gw + L[11].w_off, gw + L[11].b_off, 1, nullptr, false));
    unsigned int* gm = (u->track_amax && deferred) ? u->amax + 12 : nullptr;       // gm[l]: max|output gradient of layer l| over the sweep
    SOL_TRY(launch_conv5x5(st, B, Y, X, 2, 32, g_corr, wT + L[11].w_off, nullptr, nullptr, s.acts[10], SOL_ACT_DLRELU, a, gS, gm ? gm + 10 : nullptr, true));
    for (int k = 5; k >= 1; --k) {
        const LayerDesc& l1 = L[2 * k - 1];
        const LayerDesc& l2 = L[2 * k];
        const float* a_prev = s.acts[2 * k - 2];
        const float* t_k = s.acts[2 * k - 1];
        // pick scratch buffers that do not alias gS (non-deferred mode rotates three buffers)
        float* fT = spare[0] == gS ? spare[1] : spare[0];
        float* gT = gout(2 * k - 1, fT);
        float* fN = (spare[0] != gS && spare[0] != gT) ? spare[0] : ((spare[1] != gS && spare[1] != gT) ? spare[1] : spare[2]);
        float* gN = (k >= 2) ? gout(2 * k - 2, fN) : (deferred ? u->g0_st + (size_t)step * u->nA : fN);
        // gS = d/d(a_{k-1} + conv_{2k}(t_k) + b)
        if (!deferred) {
            // (X % 4 != 0: no partial-sum kernel for such rows, the generic kernel accumulates into gw directly — zeroed at the start of the sweep)
            if (X % 4 == 0) SOL_TRY(launch_wgrad(st, B, Y, X, 32, 32, t_k, gS, nullptr, nullptr, !first, u->partials + u->partial_stride * (2 * k - 1), false));
            else SOL_TRY(launch_wgrad(st, B, Y, X, 32, 32, t_k, gS, gw + l2.w_off, gw + l2.b_off, 1, nullptr, false));
                }
        const float* p2 = tc ? u->wprep_bwd + tc_weights_floats() * (2 * k - 1) : nullptr;
        const float* p1 = tc ? u->wprep_bwd + tc_weights_floats() * (2 * k - 2) : nullptr;
        SOL_TRY(layer_conv(u, st, gS, wT + l2.w_off, p2, nullptr, nullptr, t_k, SOL_ACT_DLRELU, a, gT, gm ? gm + 2 * k - 1 : nullptr));
        if (!deferred) {
            if (X % 4 == 0) SOL_TRY(launch_wgrad(st, B, Y, X, 32, 32, a_prev, gT, nullptr, nullptr, !first, u->partials + u->partial_stride * (2 * k - 2), false));
            else SOL_TRY(launch_wgrad(st, B, Y, X, 32, 32, a_prev, gT, gw + l1.w_off, gw + l1.b_off, 1, nullptr, false));
                }
        SOL_TRY(layer_conv(u, st, gT, wT + l1.w_off, p1, nullptr, gS, a_prev, SOL_ACT_DLRELU, a, gN, (gm && k >= 2) ? gm + 2 * k - 2 : nullptr));
        gS = gN;
    }
    // input layer (cin0 -> 32): gS is the gradient w.r.t. its pre-activation
    if (!deferred)
        SOL_TRY(launch_wgrad(st, B, Y, X, L[0].cin, 32, s.feat, gS, gw + L[0].w_off, gw + L[0].b_off, 1, nullptr, false));
    return launch_conv5x5(st, B, Y, X, 32, L[0].cin, gS, wT + L[0].w_off, nullptr, nullptr, nullptr, SOL_ACT_NONE, a, g_feat, nullptr, true);
}

// nsteps_override > 0: forward-only rollout of that many frames which recycles the stash of step 0 (no adjoint afterwards)
int do_forward(sol_unroll* u, cudaStream_t st, const float* weights, const float* re, const float* rho0, const float* vy0, const float* vx0,
               const float* gt_vy, const float* gt_vx, float* loss_steps, float* pred_vy, float* pred_vx, float* pred_rho,
               int nsteps_override = 0) {
    sol_plan* p = u->plan;
    const sol_unroll_cfg& c = u->cfg;
    const bool ring = nsteps_override > 0;
    const int B = c.B, m = ring ? nsteps_override : c.msteps;
    const size_t NY = p->NY() * B, NX = p->NX() * B, NC = p->NC() * B;
    const bool dens = c.with_density && rho0 != nullptr;
    const bool burgers = p->boundary == SOL_BOUNDARY_PERIODIC;
    if (burgers) {
        SOL_CHECK(u->burgers_set, "unroll: PERIODIC plan without sol_unroll_set_burgers()");
        SOL_CHECK(!dens, "unroll: the burgers scene has no marker density");
    } else {
        SOL_CHECK(re != nullptr, "unroll: Re[B] required for the karman scene");
        SOL_CUDA(cudaMemcpyAsync(u->re_buf, re, sizeof(float) * B, cudaMemcpyDeviceToDevice, st));
        re = u->re_buf;
    }
    if (gt_vy) SOL_CUDA(cudaMemsetAsync(loss_steps, 0, sizeof(float) * m, st));
    const bool det = sol::g_deterministic != 0 && !ring;
    if (det && gt_vy) SOL_CUDA(cudaMemsetAsync(u->loss_part, 0, sizeof(float) * (size_t)m * u->loss_part_stride, st));
    const bool mars = c.model == SOL_MODEL_MARS_MOON;
    if (conv_path_is_tc() && mars) {
        SOL_TRY(launch_split_weights_multi(st, weights + u->L[1].w_off, u->L[2].w_off - u->L[1].w_off, u->wprep_fwd, 10));
    }
    // operand maxima for the 3xFP16 weight-gradient GEMM: tracked by the 3xFP16 conv kernels as they produce the tensors
    u->track_amax = mars && sol::g_conv_path == 2 && sol::g_wgrad_path == 2 && (p->Y % 16 == 0) && (p->X % 8 == 0);
    SOL_CUDA(cudaMemsetAsync(u->amax, 0, 24 * sizeof(unsigned int), st));
    const float* cvy = vy0; const float* cvx = vx0; const float* crho = dens ? rho0 : nullptr;
    const bool fuse_io = sol::g_fuse_solver_io && cg_fuses(p, B);
    for (int i = 0; i < m; ++i) {
        StepStash& s = u->stash[ring ? 0 : i];
        int* it_slot = u->iters + (size_t)(ring ? 0 : i) * B;
        float* nvy = pred_vy ? pred_vy + (size_t)i * NY : ((i & 1) ? u->sB_vy : u->sA_vy);
        float* nvx = pred_vx ? pred_vx + (size_t)i * NX : ((i & 1) ? u->sB_vx : u->sA_vx);
        float* nrho = dens ? (pred_rho ? pred_rho + (size_t)i * NC : ((i & 1) ? u->rhoB : u->rhoA)) : nullptr;
        if (burgers) {
            // BurgersTest.step_with_f (burgers_train.py:178-187, 382-396): advect -> diffuse(nu*dt) -> + dt*f_i; the stash
            // keeps the step's INPUT velocity (what the advection adjoint needs)
            const float* fy = u->bf_vy ? u->bf_vy + (size_t)i * NY : nullptr;
            const float* fx = u->bf_vx ? u->bf_vx + (size_t)i * NX : nullptr;
            SOL_CUDA(cudaMemcpyAsync(s.vy1, cvy, sizeof(float) * NY, cudaMemcpyDeviceToDevice, st));
            SOL_CUDA(cudaMemcpyAsync(s.vx1, cvx, sizeof(float) * NX, cudaMemcpyDeviceToDevice, st));
            SOL_TRY(launch_advect(p, st, B, c.dt, s.vy1, s.vx1, nullptr, u->vy2, u->vx2, nullptr));
            SOL_TRY(launch_burgers_diffuse(p, st, B, u->visc * c.dt, u->bk_y, u->bk_x, u->vy2, u->vx2, fy, fx, c.dt, u->vy3, u->vx3));
            SOL_TRY(launch_to_feature_burgers(p, st, B, u->vy3, u->vx3, fy, fx, c.sig_vy, c.sig_vx, u->sig_fy, u->sig_fx, c.cin0, s.feat));
            SOL_TRY(cnn_forward(u, st, weights, s, u->corr));
            SOL_TRY(launch_correct_loss(p, st, B, u->vy3, u->vx3, u->corr, c.sig_vy, c.sig_vx, gt_vy ? gt_vy + (size_t)i * NY : nullptr,
                                        gt_vx ? gt_vx + (size_t)i * NX : nullptr, 1.0f / (float)m, nvy, nvx, s.gl_vy, s.gl_vx,
                                        gt_vy ? loss_steps + i : nullptr, (det && gt_vy) ? u->loss_part + (size_t)i * u->loss_part_stride : nullptr));
            cvy = nvy; cvx = nvx;
            continue;
        }
        NvtxRange r_step("fwd step %d", i);
        {
            NvtxRange r("diffuse+bc, advect");
            if (sol::g_fuse_stencil) {
                SOL_TRY(launch_diffuse_advect(p, st, B, re, c.dt, c.res, cvy, cvx, crho, s.vy1, s.vx1, u->vy2, u->vx2, nrho));
            } else {
                SOL_TRY(launch_diffuse_bc(p, st, B, re, c.dt, c.res, cvy, cvx, s.vy1, s.vx1));
                SOL_TRY(launch_advect(p, st, B, c.dt, s.vy1, s.vx1, crho, u->vy2, u->vx2, nrho));
            }
        }
        NvtxRange r_proj("pressure projection (+features)");
        if (fuse_io && c.cin0 == 3) {      // the projection kernel also writes the CNN features of the projected velocity
            CgFuse f; f.feat_out = s.feat; f.re = re; f.isy = 1.0f / c.sig_vy; f.isx = 1.0f / c.sig_vx; f.isr = 1.0f / c.sig_ext; f.cfeat = 3;
            SOL_TRY(launch_cg(p, st, B, 1, nullptr, nullptr, u->vy2, u->vx2, u->vy3, u->vx3, it_slot, &f));
        } else {
            SOL_TRY(launch_cg(p, st, B, 1, nullptr, nullptr, u->vy2, u->vx2, u->vy3, u->vx3, it_slot));
            SOL_TRY(launch_to_feature(p, st, B, u->vy3, u->vx3, re, c.sig_vy, c.sig_vx, c.sig_ext, s.feat));
        }
        r_proj.~NvtxRange(); r_proj.on = false;
        {
            NvtxRange r("correction cnn");
            SOL_TRY(cnn_forward(u, st, weights, s, u->corr));
        }
        NvtxRange r_loss("correct + loss");
        SOL_TRY(launch_correct_loss(p, st, B, u->vy3, u->vx3, u->corr, c.sig_vy, c.sig_vx, gt_vy ? gt_vy + (size_t)i * NY : nullptr,
                                    gt_vx ? gt_vx + (size_t)i * NX : nullptr, 1.0f / (float)m, nvy, nvx, s.gl_vy, s.gl_vx,
                                    gt_vy ? loss_steps + i : nullptr, (det && gt_vy) ? u->loss_part + (size_t)i * u->loss_part_stride : nullptr));
        cvy = nvy; cvx = nvx; crho = nrho;
    }
    if (det && gt_vy) {      // ordered sum of the per-CTA loss partials of every step (unused slots are zero)
        SOL_TRY(launch_loss_finalize(st, m, u->loss_part, u->loss_part_stride, loss_steps));
    }
    u->forward_done = !ring;
    u->last_re = u->re_buf;
    u->have_loss = !ring && gt_vy != nullptr;
    return SOL_OK;
}

int do_backward(sol_unroll* u, cudaStream_t st, const float* weights, const float* re, float* gw, float* g_vy0, float* g_vx0) {
    sol_plan* p = u->plan;
    const sol_unroll_cfg& c = u->cfg;
    const int B = c.B, m = c.msteps;
    SOL_CUDA(cudaMemsetAsync(gw, 0, sizeof(float) * u->nparams, st));
    const bool mars = c.model == SOL_MODEL_MARS_MOON;
    u->deferred_wgrad = mars && (sol::g_wgrad_path >= 2) && (p->Y % 16 == 0) && (p->X % 8 == 0);
    for (size_t l = 0; l < u->L.size(); ++l)
        SOL_TRY(launch_flip_weights(st, u->L[l].cin, u->L[l].cout, weights + u->L[l].w_off, u->wT + u->L[l].w_off));
    if (conv_path_is_tc() && mars) {
        SOL_TRY(launch_split_weights_multi(st, u->wT + u->L[1].w_off, u->L[2].w_off - u->L[1].w_off, u->wprep_bwd, 10));
    }
    // ---- deferred weight gradients: work items (layer, step range) over the stashed activations / output gradients.
    // While an adjoint pressure solve occupies B SMs for ~100 us, the other SMs are idle: items whose steps are already
    // complete run there on a side stream (option "wgrad_overlap"), the remainder after the sweep.
    struct WgItem { int layer, step0, nsteps; };
    int slots[12] = {0};      // per layer: CTA partial-sum slots that already hold data (launches may use different CTA counts)
    const size_t in_stride = (m > 1) ? (size_t)(u->stash[1].acts[0] - u->stash[0].acts[0]) : u->nA;
    const int tiles_step = (p->X / 8) * (p->Y / 16) * B;
    const bool burgers = p->boundary == SOL_BOUNDARY_PERIODIC;      // no pressure solve, hence no solve windows
    // (a) iterative solvers: the adjoint solve of a step keeps only B SMs busy for ~100 us: weight-gradient items fill that window.
    // (b) direct projection (no window worth filling): the adjoint conv chain needs at most 2 CTAs on tiles/2 SMs, so a persistent
    //     background launch on the SMs beyond that runs beside it without lengthening its critical path.
    const bool det = sol::g_deterministic != 0;       // one stream, private CTA slots for the thin layers
    const bool overlap = u->deferred_wgrad && sol::g_wgrad_overlap && !det && B + 17 <= p->sm_count && !burgers && !direct_for_batch(p, B);
    const int bg_free = p->sm_count - (tiles_step + 1) / 2;      // SMs the two-per-SM conv tiles leave alone
    const int bg_ctas = std::min(sol::g_wgrad_bg_ctas, bg_free);
    const bool background = u->deferred_wgrad && sol::g_wgrad_overlap && !det && !overlap && !burgers && bg_ctas >= 16 && m >= 2 * sol::g_wgrad_bg_chunk;
    const int sm_budget = overlap ? p->sm_count - B - 1 : p->sm_count;        // SMs left beside the solve's B CTAs
    const int nct32 = tiles_step < sm_budget ? tiles_step : sm_budget;
    if ((overlap || background) && !u->sstream) {
        SOL_CUDA(cudaStreamCreateWithFlags(&u->sstream, cudaStreamNonBlocking));
        SOL_CUDA(cudaEventCreateWithFlags(&u->ev_wfork, cudaEventDisableTiming));
        SOL_CUDA(cudaEventCreateWithFlags(&u->ev_wjoin, cudaEventDisableTiming));
    }
    auto launch_item = [&](cudaStream_t s, const WgItem& it, int ctas) -> int {
        const int l = it.layer;
        if (l == 11)
            return launch_wgrad_thin_multi(s, it.nsteps, B, p->Y, p->X, 32, 2, u->stash[it.step0].acts[10], in_stride,
                                           u->gcorr_st + (size_t)it.step0 * p->NC() * B * 2, (size_t)p->NC() * B * 2, gw + u->L[11].w_off,
                                           gw + u->L[11].b_off, 2 * ctas, det ? u->thin_part : nullptr);
        if (l == 0)
            return launch_wgrad_thin_multi(s, it.nsteps, B, p->Y, p->X, u->L[0].cin, 32, u->stash[it.step0].feat, in_stride,
                                           u->g0_st + (size_t)it.step0 * u->nA, u->nA, gw + u->L[0].w_off, gw + u->L[0].b_off, 2 * ctas,
                                           det ? u->thin_part : nullptr);
        int nctas = 0;
        const float* act_in = u->stash[it.step0].acts[l - 1];
        const float* g_out = u->gst + ((size_t)(l - 1) * m + it.step0) * u->nA;
        if (sol::g_wgrad_path == 3) {
            SOL_TRY(launch_wgrad_c32_tc(s, ctas, it.nsteps, B, p->Y, p->X, act_in, in_stride, g_out, u->nA, u->partials + u->partial_stride * (l - 1),
                                        &nctas, slots[l]));
        } else {
            if (!u->track_amax) {       // the tensors were produced by kernels that do not track their maxima: measure them now
                for (int k = 0; k < it.nsteps; ++k) {
                    SOL_TRY(launch_amax(s, act_in + (size_t)k * in_stride, u->nA, u->amax + l));
                    SOL_TRY(launch_amax(s, g_out + (size_t)k * u->nA, u->nA, u->amax + 12 + l));
                }
            }
            SOL_TRY(launch_wgrad_c32_h(s, ctas, it.nsteps, B, p->Y, p->X, act_in, in_stride, g_out, u->nA, u->amax + l, u->amax + 12 + l,
                                       u->partials + u->partial_stride * (l - 1), &nctas, slots[l]));
        }
        if (nctas > slots[l]) slots[l] = nctas;
        return SOL_OK;
    };
    // Greedy schedule: done[l] = steps >= done[l] of layer l are already issued.  In the window of step i the steps
    // >= i are complete; the layer with the largest backlog gets an item as long as its estimated time fits what is
    // left of the window (cost model in us, measured on B200: a launch costs ~15 us of prologue + partial-sum flush
    // + launch gap, a round of nct32 128-pixel tiles ~7.5 us; the thin layers ~7 us per step and 192 tiles).
    int done[12];
    for (int l = 0; l < 12; ++l) done[l] = m;
    auto item_cost = [&](int l, int n) -> float {
        if (l == 0 || l == 11) return 4.0f + 7.0f * (float)n * (float)tiles_step / 192.0f;
        return 15.0f + 7.5f * (float)(((long)tiles_step * n + nct32 - 1) / nct32);
    };
    // adjoint solve + advection adjoint on this grid (option "wgrad_window_us" overrides the 128x64 figure)
    const float window_us = (float)sol::g_wgrad_window_us * (float)(p->Y * p->X) / 8192.0f;
    auto fill_window = [&](cudaStream_t s, int i, float budget) -> int {
        float left = budget;
        bool any = false;
        for (;;) {
            // small items are mostly launch overhead: a layer waits until 3 steps (thin layers: 2) have piled up
            int best = -1, backlog = 0;
            for (int l = 0; l < 12; ++l) {
                const int bl = done[l] - i, nmin = (l == 0 || l == 11) ? 2 : 3;
                if (bl >= nmin && bl > backlog) { backlog = bl; best = l; }
            }
            if (best < 0) break;
            int n = backlog;
            while (n > 0 && item_cost(best, n) > left) --n;
            if (n == 0) {
                if (any) break;
                n = 1;                                  // always make progress: one step of the fullest layer
            }
            SOL_TRY(launch_item(s, WgItem{best, done[best] - n, n}, nct32));
            done[best] -= n;
            left -= item_cost(best, n);
            any = true;
            if (left <= 0.0f) break;
        }
        return SOL_OK;
    };

    bool bg_pending = false;
    const float* Gy = u->stash[m - 1].gl_vy;
    const float* Gx = u->stash[m - 1].gl_vx;
    const bool fuse_io = sol::g_fuse_solver_io && !burgers && cg_fuses(p, B);
    bool corr_ready = false;
    if (fuse_io) {      // the first scatter targets; every later pair is zeroed by the preceding diffusion adjoint
        float* Hy0 = ((m - 1) & 1) ? u->H2_vy : u->H_vy; float* Hx0 = ((m - 1) & 1) ? u->H2_vx : u->H_vx;
        SOL_CUDA(cudaMemsetAsync(Hy0, 0, sizeof(float) * p->NY() * B, st));
        SOL_CUDA(cudaMemsetAsync(Hx0, 0, sizeof(float) * p->NX() * B, st));
    }
    for (int i = m - 1; i >= 0; --i) {
        StepStash& s = u->stash[i];
        float* g_corr = u->deferred_wgrad ? u->gcorr_st + (size_t)i * p->NC() * B * 2 : u->g_corr;
        NvtxRange r_step("bwd step %d", i);
        if (!corr_ready) SOL_TRY(launch_corr_bwd(p, st, B, Gy, Gx, c.sig_vy, c.sig_vx, g_corr));   // else: written by diffuse_bc_bwd of step i+1
        {
            NvtxRange r("correction cnn adjoint");
            SOL_TRY(cnn_backward(u, st, weights, gw, s, g_corr, u->g_feat, i == m - 1, i));
        }
        if (background && i >= sol::g_wgrad_bg_chunk && done[1] - i >= sol::g_wgrad_bg_chunk) {
            // the output gradients of steps [i, done) are complete: their weight-gradient items run on the side stream from here on
            SOL_CUDA(cudaEventRecord(u->ev_wfork, st));
            SOL_CUDA(cudaStreamWaitEvent(u->sstream, u->ev_wfork, 0));
            for (int l = 0; l <= 11; ++l) {
                SOL_TRY(launch_item(u->sstream, WgItem{l, i, done[l] - i}, bg_ctas));
                done[l] = i;
            }
            bg_pending = true;
        }
        if (!fuse_io) SOL_TRY(launch_feat_bwd(p, st, B, Gy, Gx, u->g_feat, c.cin0, c.sig_vy, c.sig_vx, u->H_vy, u->H_vx));
        if (burgers) {
            // adjoint of advect -> diffuse (+ dt*f: no state dependence); the periodic diffusion operator is symmetric
            SOL_TRY(launch_burgers_diffuse(p, st, B, u->visc * c.dt, u->bk_y, u->bk_x, u->H_vy, u->H_vx, nullptr, nullptr, 0.f, u->K_vy, u->K_vx));
            SOL_TRY(launch_advect_bwd(p, st, B, c.dt, s.vy1, s.vx1, u->K_vy, u->K_vx, u->H_vy, u->H_vx));
            if (i > 0) {
                float* ny = u->G_vy[i & 1]; float* nx = u->G_vx[i & 1];
                float* gc_next = nullptr;
                if (sol::g_fuse_small) gc_next = u->deferred_wgrad ? u->gcorr_st + (size_t)(i - 1) * p->NC() * B * 2 : u->g_corr;
                SOL_TRY(launch_add_faces(p, st, B, u->H_vy, u->H_vx, u->stash[i - 1].gl_vy, u->stash[i - 1].gl_vx, ny, nx, gc_next, c.sig_vy,
                                         c.sig_vx));
                corr_ready = gc_next != nullptr;
                Gy = ny; Gx = nx;
            } else if (g_vy0 && g_vx0) {
                SOL_TRY(launch_add_faces(p, st, B, u->H_vy, u->H_vx, nullptr, nullptr, g_vy0, g_vx0));
            }
            continue;
        }
        bool joined = true;
        if (overlap && i > 0) {     // the items of the last window would only delay the end of the sweep: they are merged below
            SOL_CUDA(cudaEventRecord(u->ev_wfork, st));
            SOL_CUDA(cudaStreamWaitEvent(u->sstream, u->ev_wfork, 0));
            SOL_TRY(fill_window(u->sstream, i, window_us));
            SOL_CUDA(cudaEventRecord(u->ev_wjoin, u->sstream));
            joined = false;
        }
        // scatter targets of the advection adjoint: with the fused solver I/O nothing else uses H, so two buffer pairs alternate and
        // the diffusion adjoint of a step zeroes the pair of the NEXT one (no memset nodes inside the sweep)
        float* Hy = (fuse_io && (i & 1)) ? u->H2_vy : u->H_vy;
        float* Hx = (fuse_io && (i & 1)) ? u->H2_vx : u->H_vx;
        float* Zy = fuse_io ? ((i & 1) ? u->H_vy : u->H2_vy) : nullptr;
        float* Zx = fuse_io ? ((i & 1) ? u->H_vx : u->H2_vx) : nullptr;
        if (fuse_io) {      // the adjoint projection adds the feature gradient to the incoming velocity gradient itself
            CgFuse f; f.gfeat_in = u->g_feat; f.isy = 1.0f / c.sig_vy; f.isx = 1.0f / c.sig_vx; f.cfeat = c.cin0;
            SOL_TRY(launch_cg(p, st, B, 1, nullptr, nullptr, Gy, Gx, u->K_vy, u->K_vx, u->iters + (size_t)(m + i) * B, &f));
        } else {
            SOL_TRY(launch_cg(p, st, B, 1, nullptr, nullptr, u->H_vy, u->H_vx, u->K_vy, u->K_vx, u->iters + (size_t)(m + i) * B));
        }
        SOL_TRY(launch_advect_bwd(p, st, B, c.dt, s.vy1, s.vx1, u->K_vy, u->K_vx, Hy, Hx, fuse_io));
        if (!joined) SOL_CUDA(cudaStreamWaitEvent(st, u->ev_wjoin, 0));
        if (i > 0) {
            float* ny = u->G_vy[i & 1]; float* nx = u->G_vx[i & 1];
            float* gc_next = nullptr;       // fused corr_bwd of step i-1
            if (sol::g_fuse_small) gc_next = u->deferred_wgrad ? u->gcorr_st + (size_t)(i - 1) * p->NC() * B * 2 : u->g_corr;
            SOL_TRY(launch_diffuse_bc_bwd(p, st, B, re, c.dt, c.res, Hy, Hx, ny, nx, u->stash[i - 1].gl_vy, u->stash[i - 1].gl_vx,
                                          gc_next, c.sig_vy, c.sig_vx, Zy, Zx));
            corr_ready = gc_next != nullptr;
            Gy = ny; Gx = nx;
        } else if (g_vy0 && g_vx0) {
            SOL_TRY(launch_diffuse_bc_bwd(p, st, B, re, c.dt, c.res, Hy, Hx, g_vy0, g_vx0, nullptr, nullptr));
        }
    }
    if (u->deferred_wgrad) {
        // what the solve windows did not absorb: ONE launch per layer over its remaining steps [0, done[l]), then the
        // per-layer reduction of the CTA partial sums
        NvtxRange r_w("deferred weight gradients (tail + finalize)");
        if (bg_pending) {
            SOL_CUDA(cudaEventRecord(u->ev_wjoin, u->sstream));
            SOL_CUDA(cudaStreamWaitEvent(st, u->ev_wjoin, 0));
        }
        for (int l = 0; l <= 11; ++l)
            if (done[l] > 0) { SOL_TRY(launch_item(st, WgItem{l, 0, done[l]}, nct32)); done[l] = 0; }
        // the ten hidden layers' [weights | bias] blocks are adjacent in the Keras-ordered buffer: one reduction launch
        bool adjacent = true;
        for (int l = 1; l <= 10; ++l)
            adjacent &= u->L[l].b_off == u->L[l].w_off + 25 * 32 * 32 && (l == 10 || u->L[l + 1].w_off == u->L[l].b_off + 32);
        if (adjacent) {
            SOL_TRY(launch_wgrad_finalize_multi(st, slots + 1, u->partials, u->partial_stride, gw + u->L[1].w_off));
        } else {
            for (int l = 1; l <= 10; ++l)
                SOL_TRY(launch_wgrad_finalize_n(st, slots[l], u->partials + u->partial_stride * (l - 1), gw + u->L[l].w_off, gw + u->L[l].b_off, 0));
        }
        return SOL_OK;
    }
    if (!mars || p->X % 4 != 0) return SOL_OK;      // accumulated into gw step by step
    for (int l = 1; l <= 10; ++l)
        SOL_TRY(launch_wgrad(st, B, p->Y, p->X, 32, 32, nullptr, nullptr, gw + u->L[l].w_off, gw + u->L[l].b_off, 0,
                             u->partials + u->partial_stride * (l - 1), true));
    return SOL_OK;
}

}  // namespace

extern "C" size_t sol_unroll_workspace_bytes(const sol_plan* plan, const sol_unroll_cfg* cfg) {
    if (check_cfg(plan, cfg) != SOL_OK) return 0;
    sol_unroll tmp;
    tmp.plan = const_cast<sol_plan*>(plan);
    tmp.cfg = *cfg;
    if (build_layers(cfg->model, cfg->cin0, tmp.L) != SOL_OK) return 0;
    tmp.nparams = param_count(tmp.L);
    size_t total = 0;
    carve(&tmp, nullptr, &total);
    return total;
}

extern "C" int sol_unroll_create(sol_plan* plan, const sol_unroll_cfg* cfg, void* workspace, size_t workspace_bytes, sol_unroll** out) {
    SOL_CHECK(out != nullptr && workspace != nullptr, "sol_unroll_create: NULL pointer");
    SOL_TRY(check_cfg(plan, cfg));
    sol_unroll* u = new (std::nothrow) sol_unroll();
    if (!u) return fail(SOL_ERR_INVALID, "out of host memory");
    u->plan = plan;
    u->cfg = *cfg;
    build_layers(cfg->model, cfg->cin0, u->L);
    u->nparams = param_count(u->L);
    size_t total = 0;
    carve(u, workspace, &total);
    if (total > workspace_bytes) { delete u; return fail(SOL_ERR_WORKSPACE, "sol_unroll_create: workspace too small (see sol_unroll_workspace_bytes)"); }
    if (((uintptr_t)workspace & 255) != 0) { delete u; return fail(SOL_ERR_INVALID, "sol_unroll_create: workspace must be 256-byte aligned"); }
    *out = u;
    return SOL_OK;
}

extern "C" int sol_unroll_destroy(sol_unroll* u) {
    if (!u) return SOL_OK;
    if (u->gexec) cudaGraphExecDestroy(u->gexec);
    if (u->ev_fork) cudaEventDestroy(u->ev_fork);
    if (u->ev_join) cudaEventDestroy(u->ev_join);
    if (u->gstream) cudaStreamDestroy(u->gstream);
    if (u->ev_wfork) cudaEventDestroy(u->ev_wfork);
    if (u->ev_wjoin) cudaEventDestroy(u->ev_wjoin);
    if (u->sstream) cudaStreamDestroy(u->sstream);
    delete u;
    return SOL_OK;
}

extern "C" int sol_unroll_forward(sol_unroll* u, void* stream, const float* weights, const float* re, const float* rho0, const float* vy0,
                                  const float* vx0, const float* gt_vy, const float* gt_vx, float* loss_steps, float* pred_vy,
                                  float* pred_vx, float* pred_rho) {
    SOL_CHECK(u && weights && vy0 && vx0, "sol_unroll_forward: NULL pointer");
    SOL_CHECK((gt_vy == nullptr) == (gt_vx == nullptr), "sol_unroll_forward: gt_vy and gt_vx go together");
    SOL_CHECK(!gt_vy || loss_steps, "sol_unroll_forward: loss_steps required with ground truth");
    SOL_CHECK((pred_vy == nullptr) == (pred_vx == nullptr), "sol_unroll_forward: pred_vy and pred_vx go together");
    return do_forward(u, (cudaStream_t)stream, weights, re, rho0, vy0, vx0, gt_vy, gt_vx, loss_steps, pred_vy, pred_vx, pred_rho);
}

extern "C" int sol_unroll_rollout(sol_unroll* u, void* stream, const float* weights, const float* re, const float* rho0, const float* vy0,
                                  const float* vx0, int nsteps, float* pred_vy, float* pred_vx, float* pred_rho) {
    SOL_CHECK(u && weights && vy0 && vx0 && pred_vy && pred_vx && nsteps >= 1, "sol_unroll_rollout: bad arguments");
    SOL_CHECK(pred_rho == nullptr || (rho0 != nullptr && u->cfg.with_density), "sol_unroll_rollout: pred_rho needs rho0 and cfg.with_density");
    return do_forward(u, (cudaStream_t)stream, weights, re, rho0, vy0, vx0, nullptr, nullptr, nullptr, pred_vy, pred_vx, pred_rho, nsteps);
}

extern "C" int sol_unroll_backward(sol_unroll* u, void* stream, const float* weights, float* grad_weights, float* g_vy0, float* g_vx0) {
    SOL_CHECK(u && weights && grad_weights, "sol_unroll_backward: NULL pointer");
    SOL_CHECK(u->forward_done && u->have_loss, "sol_unroll_backward: call sol_unroll_forward with ground truth first");
    return do_backward(u, (cudaStream_t)stream, weights, u->last_re, grad_weights, g_vy0, g_vx0);
}

extern "C" int sol_unroll_train_iter(sol_unroll* u, void* stream, const float* weights, const float* re, const float* rho0, const float* vy0,
                                     const float* vx0, const float* gt_vy, const float* gt_vx, float* loss_steps, float* grad_weights) {
    SOL_CHECK(u && weights && vy0 && vx0 && gt_vy && gt_vx && loss_steps && grad_weights, "sol_unroll_train_iter: NULL pointer");
    cudaStream_t caller = (cudaStream_t)stream;
    const void* key[13] = {weights, re, rho0, vy0, vx0, gt_vy, gt_vx, loss_steps, grad_weights, u->bk_y, u->bk_x, u->bf_vy, u->bf_vx};
    if (!u->cfg.use_graph) {
        SOL_TRY(do_forward(u, caller, weights, re, rho0, vy0, vx0, gt_vy, gt_vx, loss_steps, nullptr, nullptr, nullptr));
        return do_backward(u, caller, weights, u->re_buf, grad_weights, nullptr, nullptr);
    }
    if (!u->gstream) {
        SOL_CUDA(cudaStreamCreateWithFlags(&u->gstream, cudaStreamNonBlocking));
        SOL_CUDA(cudaEventCreateWithFlags(&u->ev_fork, cudaEventDisableTiming));
        SOL_CUDA(cudaEventCreateWithFlags(&u->ev_join, cudaEventDisableTiming));
    }
    cudaStream_t st = u->gstream;
    auto run = [&]() -> int {
        SOL_TRY(do_forward(u, st, weights, re, rho0, vy0, vx0, gt_vy, gt_vx, loss_steps, nullptr, nullptr, nullptr));
        return do_backward(u, st, weights, u->re_buf, grad_weights, nullptr, nullptr);
    };
    auto join = [&]() -> int {
        SOL_CUDA(cudaEventRecord(u->ev_join, st));
        SOL_CUDA(cudaStreamWaitEvent(caller, u->ev_join, 0));
        return SOL_OK;
    };
    SOL_CUDA(cudaEventRecord(u->ev_fork, caller));
    SOL_CUDA(cudaStreamWaitEvent(st, u->ev_fork, 0));
    const unsigned long long epoch = g_cfg_epoch.load(std::memory_order_relaxed);
    const bool same = memcmp(key, u->gkey, sizeof(key)) == 0 && epoch == u->gepoch;
    if (same && u->gexec) {
        SOL_CUDA(cudaGraphLaunch(u->gexec, st));
        sol::g_launches.fetch_add(u->graph_kernels, std::memory_order_relaxed);
        return join();
    }
    if (!same || u->warm == 0) {
        // first call with these buffers: run eagerly (also sets every kernel attribute outside capture)
        if (u->gexec) { cudaGraphExecDestroy(u->gexec); u->gexec = nullptr; }
        memcpy(u->gkey, key, sizeof(key));
        u->gepoch = epoch;
        u->warm = 1;
        SOL_TRY(run());
        return join();
    }
    // second call with the same buffers: capture, instantiate, launch
    cudaGraph_t graph = nullptr;
    SOL_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    const unsigned long long before = sol::g_launches.load();
    int rc = run();
    cudaError_t ce = cudaStreamEndCapture(st, &graph);
    u->graph_kernels = sol::g_launches.load() - before;   // kernel nodes in the graph
    sol::g_launches.store(before);                        // captured launches did not execute
    if (rc != SOL_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) { snprintf(sol::g_err, sizeof(sol::g_err), "cudaStreamEndCapture: %s", cudaGetErrorString(ce)); return SOL_ERR_CUDA; }
    ce = cudaGraphInstantiate(&u->gexec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) { u->gexec = nullptr; snprintf(sol::g_err, sizeof(sol::g_err), "cudaGraphInstantiate: %s", cudaGetErrorString(ce)); return SOL_ERR_CUDA; }
    SOL_CUDA(cudaGraphLaunch(u->gexec, st));
    sol::g_launches.fetch_add(u->graph_kernels, std::memory_order_relaxed);
    return join();
}

extern "C" int sol_unroll_set_burgers(sol_unroll* u, float viscosity, const float* diff_kernel_y, const float* diff_kernel_x,
                                      const float* f_vy, const float* f_vx, float sig_fy, float sig_fx) {
    SOL_CHECK(u != nullptr, "sol_unroll_set_burgers: NULL unroll");
    SOL_CHECK(u->plan->boundary == SOL_BOUNDARY_PERIODIC, "sol_unroll_set_burgers: PERIODIC plans only");
    SOL_CHECK((diff_kernel_y == nullptr) == (diff_kernel_x == nullptr) && (f_vy == nullptr) == (f_vx == nullptr),
              "sol_unroll_set_burgers: kernels / forces come in pairs");
    SOL_CHECK(viscosity >= 0.f, "sol_unroll_set_burgers: negative viscosity");
    if (u->cfg.cin0 == 4) SOL_CHECK(f_vy != nullptr && sig_fy > 0.f && sig_fx > 0.f, "sol_unroll_set_burgers: 4 feature channels need forces and positive force sigmas");
    u->visc = viscosity; u->bk_y = diff_kernel_y; u->bk_x = diff_kernel_x; u->bf_vy = f_vy; u->bf_vx = f_vx;
    u->sig_fy = sig_fy; u->sig_fx = sig_fx;
    u->burgers_set = true;
    cfg_changed();
    return SOL_OK;
}

extern "C" int sol_unroll_cg_iters(sol_unroll* u, const int** dev_iters, int* count) {
    SOL_CHECK(u && dev_iters && count, "sol_unroll_cg_iters: NULL pointer");
    *dev_iters = u->iters;
    *count = 2 * u->cfg.msteps * u->cfg.B;
    return SOL_OK;
}
