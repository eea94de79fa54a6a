// Persistent conjugate-gradient pressure projection for sm_100a.
//
// Replaces PhiFlow's divergence_free()/SparseCG (reference call sites karman_train.py:167-168,185;
// SURVEY.md §8a rows a8-a11): hard-BC face masking, staggered divergence, CG on the 5-point
// Laplace with Dirichlet p=0 outside (OPEN) and Neumann at obstacle faces, gradient subtract —
// ONE kernel launch per projection.
//
// B200 mapping: one CTA (or one thread-block cluster of CL CTAs, split along y) per simulation.
// Each thread owns a column strip of R cells; x, r, p of all CG vectors live in REGISTERS for the
// whole solve, the search direction p is exchanged through a shared-memory tile with a 1-cell
// halo (cluster: boundary rows are pushed into the neighbour CTA's halo through DSMEM), the two
// dot products per iteration are warp-shuffle + shared-memory reductions (cluster: DSMEM slots).
// HBM traffic is therefore the compulsory minimum: read v (8 B/cell), write v (8 B/cell) and
// optionally p (4 B/cell), independent of the iteration count K; the ALGORITHMIC traffic a
// streaming CG would need is (40 K + 8) B/cell, which is what bench.py reports
// against the HBM roofline.
//
// The recurrences are standard CG (x=0; r=p=d; a=rr/(p.Ap); x+=a p; r-=a Ap; b=rr'/rr; p=r+b p),
// algebraically identical to the reference's (a=(p.r)/(p.q), b=(r.q)/(p.q)) with the same
// max|r| < tol stop rule, applied per simulation.
#include <cooperative_groups.h>

#include "sol_cells.cuh"
#include "sol_internal.cuh"

SOL_TRACE_TU()

namespace cg = cooperative_groups;

namespace sol {

struct CgArgs {
    int Y, X, B;
    const float* diag;            // [Y*X]
    const unsigned char* active;  // [Y*X]
    const float* my;              // [(Y+1)*X]
    const float* mx;              // [Y*(X+1)]
    const float* rhs;             // mode 0: [B,Y,X]
    float* p_out;                 // [B,Y,X] or null
    const float* vy_in;           // mode 1
    const float* vx_in;
    float* vy_out;
    float* vx_out;
    float tol_abs, tol_rel;
    int max_it;
    int* iters;                   // [B] or null
};

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// max of NON-NEGATIVE floats in one REDUX instruction: for x >= 0 the IEEE bit patterns are ordered
__device__ __forceinline__ float wmax_nonneg(float v) {
    return __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(v)));
}

// Block/cluster-wide reduction of a sum s (and, if WITH_MAX, of a non-negative max m).  Every thread
// of every CTA of the cluster returns bit-identical results (xor butterflies are symmetric; partials
// are combined in a fixed order).  wpart: [2][2][32], clpart: [2][2][8] shared floats,
// par = reduction parity (double buffering: a buffer is rewritten only two reductions later, and a
// barrier separates every pair of reductions).
template <int CL, bool WITH_MAX>
__device__ __forceinline__ void reduce_sm(float& s, float& m, float* wpart, float* clpart, int par, int warp, int lane,
                                          int nwarps, int rank) {
    s = wsum(s);
    if (WITH_MAX) m = wmax_nonneg(m);
    float* ws = wpart + (par * 2 + 0) * 32;
    float* wm = wpart + (par * 2 + 1) * 32;
    if (lane == 0) { ws[warp] = s; if (WITH_MAX) wm[warp] = m; }
    __syncthreads();
    s = (lane < nwarps) ? ws[lane] : 0.0f;
    s = wsum(s);
    if (WITH_MAX) { m = (lane < nwarps) ? wm[lane] : 0.0f; m = wmax_nonneg(m); }
    if (CL > 1) {
        cg::cluster_group cluster = cg::this_cluster();
        float* cs = clpart + (par * 2 + 0) * 8;
        float* cm = clpart + (par * 2 + 1) * 8;
        if (warp == 0 && lane < CL) {
            float* rs = cluster.map_shared_rank(cs, lane);
            rs[rank] = s;
            if (WITH_MAX) { float* rm = cluster.map_shared_rank(cm, lane); rm[rank] = m; }
        }
        cluster.sync();
        float S = 0.0f, M = 0.0f;
#pragma unroll
        for (int c = 0; c < CL; ++c) { S += cs[c]; if (WITH_MAX) M = fmaxf(M, cm[c]); }
        s = S; m = M;
    }
}

// MODE 0: solve A p = rhs.  MODE 1: fused projection of a velocity field.
// X (cells per row, = threads per row), R (rows per thread) and NT (max threads per CTA) are
// compile-time so that every shared-memory address in the iteration loop is base + immediate.
template <int X, int R, int CL, int MODE, int NT>
__global__ void __launch_bounds__(NT, 1) k_cg(const CgArgs a) {
    pdl_sync();
    extern __shared__ float smem[];
    constexpr int PITCH = X + 2;
    const int Y = a.Y;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int TY = blockDim.y;
    const int tid = ty * X + tx;
    const int nthreads = X * TY;
    const int nwarps = nthreads >> 5;
    const int lane = tid & 31, warp = tid >> 5;
    int rank = 0;
    if (CL > 1) rank = (int)cg::this_cluster().block_rank();
    const int b = blockIdx.y;
    const int rows_cta = TY * R;
    float* ps = smem;                              // (rows_cta+2) * PITCH, halo ring of zeros
    float* wpart = ps + (rows_cta + 2) * PITCH;    // 2*2*32
    float* clpart = wpart + 128;                   // 2*2*8

    for (int k = tid; k < (rows_cta + 2) * PITCH; k += nthreads) ps[k] = 0.0f;

    const int lr0 = ty * R;                // first local row of this thread's strip
    const int j0 = rank * rows_cta + lr0;  // first global row
    const size_t NC = (size_t)Y * X, NY = (size_t)(Y + 1) * X, NX = (size_t)Y * (X + 1);
    float* const pc = ps + (lr0 + 1) * PITCH + tx + 1;   // this thread's first cell in the tile

    // Almost every cell is "regular" (fluid, 4 accessible neighbours: OPEN borders count as
    // accessible).  Warps whose cells are all regular run q = nb - 4 p with no per-cell data in
    // registers; the few warps next to the obstacle re-read diag/active (L1-resident) instead.
    float x[R], r[R], p[R];
    unsigned act = 0u;
    bool regular = true;
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const int c = (j0 + k) * X + tx;
        const bool ak = a.active[c] != 0;
        if (ak) act |= 1u << k;
        regular = regular && ak && (a.diag[c] == 4.0f);
        x[k] = 0.0f;
    }
    regular = __all_sync(0xffffffffu, regular);     // warp-uniform fast path
    if (MODE == 1) {
        const float* vy = a.vy_in + (size_t)b * NY;
        const float* vx = a.vx_in + (size_t)b * NX;
        float vlo = a.my[j0 * X + tx] * vy[j0 * X + tx];
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int j = j0 + k;
            const float vhi = a.my[(j + 1) * X + tx] * vy[(j + 1) * X + tx];
            const float xl = a.mx[j * (X + 1) + tx] * vx[j * (X + 1) + tx];
            const float xr = a.mx[j * (X + 1) + tx + 1] * vx[j * (X + 1) + tx + 1];
            r[k] = (vhi - vlo) + (xr - xl);
            vlo = vhi;
        }
    } else {
        const float* rhs = a.rhs + (size_t)b * NC;
#pragma unroll
        for (int k = 0; k < R; ++k) r[k] = ((act >> k) & 1u) ? rhs[(j0 + k) * X + tx] : 0.0f;
    }

    float rr = 0.0f, rmax = 0.0f;
#pragma unroll
    for (int k = 0; k < R; ++k) { p[k] = r[k]; rr = fmaf(r[k], r[k], rr); rmax = fmaxf(rmax, fabsf(r[k])); }
    int par = 0;
    __syncthreads();   // ps zero-fill complete before anyone writes p into it
    if (CL > 1) cg::this_cluster().sync();   // ... including remote halo pushes
    reduce_sm<CL, true>(rr, rmax, wpart, clpart, par, warp, lane, nwarps, rank); par ^= 1;
    const float tol = fmaxf(a.tol_abs, a.tol_rel * rmax);

    int it = 0;
    while (it < a.max_it && rmax > 0.0f && rmax >= tol) {
        // ---- publish p (own tile + halo rows of the neighbouring CTAs) ----
#pragma unroll
        for (int k = 0; k < R; ++k) pc[k * PITCH] = p[k];
        if (CL > 1) {
            cg::cluster_group cluster = cg::this_cluster();
            if (ty == 0 && rank > 0) {
                float* rp = cluster.map_shared_rank(ps, rank - 1);
                rp[(rows_cta + 1) * PITCH + tx + 1] = p[0];
            }
            if (ty == TY - 1 && rank < CL - 1) {
                float* rp = cluster.map_shared_rank(ps, rank + 1);
                rp[0 * PITCH + tx + 1] = p[R - 1];
            }
            cluster.sync();
        } else {
            __syncthreads();
        }
        // ---- q = A p, pq = p.q ----
        float q[R];
        float pq = 0.0f, dummy = 0.0f;
        if (regular) {
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const float up = (k + 1 < R) ? p[k + 1] : pc[(k + 1) * PITCH];
                const float dn = (k > 0) ? p[k - 1] : pc[(k - 1) * PITCH];
                const float nb = (up + dn) + (pc[k * PITCH - 1] + pc[k * PITCH + 1]);
                q[k] = fmaf(-4.0f, p[k], nb);
                pq = fmaf(p[k], q[k], pq);
            }
        } else {
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const float up = (k + 1 < R) ? p[k + 1] : pc[(k + 1) * PITCH];
                const float dn = (k > 0) ? p[k - 1] : pc[(k - 1) * PITCH];
                const float nb = (up + dn) + (pc[k * PITCH - 1] + pc[k * PITCH + 1]);
                q[k] = ((act >> k) & 1u) ? fmaf(-__ldg(a.diag + (j0 + k) * X + tx), p[k], nb) : 0.0f;
                pq = fmaf(p[k], q[k], pq);
            }
        }
        reduce_sm<CL, false>(pq, dummy, wpart, clpart, par, warp, lane, nwarps, rank); par ^= 1;
        const float alpha = (pq != 0.0f) ? __fdividef(rr, pq) : 0.0f;
        float rr_new = 0.0f;
        rmax = 0.0f;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            x[k] = fmaf(alpha, p[k], x[k]);
            r[k] = fmaf(-alpha, q[k], r[k]);
            rr_new = fmaf(r[k], r[k], rr_new);
            rmax = fmaxf(rmax, fabsf(r[k]));
        }
        reduce_sm<CL, true>(rr_new, rmax, wpart, clpart, par, warp, lane, nwarps, rank); par ^= 1;
        const float beta = (rr != 0.0f) ? __fdividef(rr_new, rr) : 0.0f;
        rr = rr_new;
#pragma unroll
        for (int k = 0; k < R; ++k) p[k] = fmaf(beta, p[k], r[k]);
        ++it;
    }

    if (a.iters && tid == 0 && rank == 0) a.iters[b] = it;

    if (MODE == 0) {
        const float* rhs = a.rhs + (size_t)b * NC;
        float* po = a.p_out + (size_t)b * NC;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int c = (j0 + k) * X + tx;
            // solid rows of the reference matrix decouple: -diag * p = rhs
            po[c] = ((act >> k) & 1u) ? x[k] : -rhs[c] / a.diag[c];
        }
        if (CL > 1) cg::this_cluster().sync();   // keep smem alive until all remote traffic is done
        return;
    }

    // ---- MODE 1: publish the pressure and subtract its masked gradient ----
#pragma unroll
    for (int k = 0; k < R; ++k) pc[k * PITCH] = x[k];
    if (CL > 1) {
        cg::cluster_group cluster = cg::this_cluster();
        if (ty == TY - 1 && rank < CL - 1) {
            float* rp = cluster.map_shared_rank(ps, rank + 1);
            rp[0 * PITCH + tx + 1] = x[R - 1];
        }
        cluster.sync();
    } else {
        __syncthreads();
    }
    {
        const float* vy = a.vy_in + (size_t)b * NY;
        const float* vx = a.vx_in + (size_t)b * NX;
        float* vyo = a.vy_out + (size_t)b * NY;
        float* vxo = a.vx_out + (size_t)b * NX;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int j = j0 + k;
            const float pdn = (k > 0) ? x[k - 1] : pc[(k - 1) * PITCH];   // p[j-1,i]; zero halo at j = 0
            const float plf = pc[k * PITCH - 1];                          // p[j,i-1]; zero halo at i = 0
            vyo[j * X + tx] = a.my[j * X + tx] * (vy[j * X + tx] - (x[k] - pdn));
            vxo[j * (X + 1) + tx] = a.mx[j * (X + 1) + tx] * (vx[j * (X + 1) + tx] - (x[k] - plf));
            if (tx == X - 1) vxo[j * (X + 1) + X] = a.mx[j * (X + 1) + X] * (vx[j * (X + 1) + X] + x[k]);
        }
        if (j0 + R == Y) vyo[Y * X + tx] = a.my[Y * X + tx] * (vy[Y * X + tx] + x[R - 1]);
        if (a.p_out) {
            float* po = a.p_out + (size_t)b * NC;
#pragma unroll
            for (int k = 0; k < R; ++k) po[(j0 + k) * X + tx] = x[k];
        }
    }
    if (CL > 1) cg::this_cluster().sync();
}

// ------------------------------------------------------------------------------------------------
template <int X, int R, int CL, int MODE, int NT>
static int launch_cg_t(const CgArgs& a, cudaStream_t st, int TY) {
    const int rows_cta = TY * R;
    const size_t smem = ((size_t)(rows_cta + 2) * (X + 2) + 128 + 32) * sizeof(float);
    auto kern = k_cg<X, R, CL, MODE, NT>;
    static size_t attr_smem = 48 * 1024;
    if (smem > attr_smem) {
        SOL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CL, a.B, 1);
    cfg.blockDim = dim3(X, TY, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    if (CL > 1) {
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CL;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    if (g_trace_names) trace_record_launch((const void*)kern);
    SOL_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
    SOL_LAUNCHED();
    return SOL_OK;
}


// ------------------------------------------------------------------------------------------------
// ANY grid size (fallback for X not in {32, 64, 128} or Y not tiling the register kernels): the same recurrences and stop rule, one
// CTA of 1024 threads per simulation, x / r / p / q in a global scratch (L2-resident), cells strided over the threads.  Not a fast
// path: it exists so that every resolution the reference's scripts accept (karman.py -r, burgers -l) runs through the same ABI.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_sum_max(float& s, float& m, float* red, int tid) {
    s = wsum(s); m = wmax_nonneg(m);
    __syncthreads();                                  // previous use of `red` is over
    if ((tid & 31) == 0) { red[tid >> 5] = s; red[32 + (tid >> 5)] = m; }
    __syncthreads();
    s = red[tid & 31]; m = red[32 + (tid & 31)];      // 32 warps
    s = wsum(s); m = wmax_nonneg(m);
}

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_cg_any(const CgArgs a, float* __restrict__ scratch) {
    pdl_sync();
    __shared__ float red[64];
    const int Y = a.Y, X = a.X, N = Y * X, tid = threadIdx.x, b = blockIdx.x;
    const size_t NY = (size_t)(Y + 1) * X, NX = (size_t)Y * (X + 1);
    float* x = scratch + (size_t)b * 4 * N;
    float* r = x + N; float* p = r + N; float* q = p + N;
    const float* vy = MODE == 1 ? a.vy_in + (size_t)b * NY : nullptr;
    const float* vx = MODE == 1 ? a.vx_in + (size_t)b * NX : nullptr;
    const float* rhs = MODE == 0 ? a.rhs + (size_t)b * N : nullptr;
    float rr = 0.0f, rmax = 0.0f;
    for (int c = tid; c < N; c += 1024) {
        const int j = c / X, i = c - j * X;
        float d;
        if (MODE == 1) {
            d = (a.my[(j + 1) * X + i] * vy[(j + 1) * X + i] - a.my[j * X + i] * vy[j * X + i]) +
                (a.mx[j * (X + 1) + i + 1] * vx[j * (X + 1) + i + 1] - a.mx[j * (X + 1) + i] * vx[j * (X + 1) + i]);
        } else {
            d = a.active[c] ? rhs[c] : 0.0f;
        }
        x[c] = 0.0f; r[c] = d; p[c] = d;
        rr = fmaf(d, d, rr); rmax = fmaxf(rmax, fabsf(d));
    }
    block_sum_max(rr, rmax, red, tid);
    const float tol = fmaxf(a.tol_abs, a.tol_rel * rmax);
    int it = 0;
    while (it < a.max_it && rmax > 0.0f && rmax >= tol) {
        __syncthreads();                              // p of every thread is in place
        float pq = 0.0f, dummy = 0.0f;
        for (int c = tid; c < N; c += 1024) {
            const int j = c / X, i = c - j * X;
            const float nb = ((j > 0 ? p[c - X] : 0.0f) + (j + 1 < Y ? p[c + X] : 0.0f)) + ((i > 0 ? p[c - 1] : 0.0f) + (i + 1 < X ? p[c + 1] : 0.0f));
            const float pc = p[c];
            const float qc = a.active[c] ? fmaf(-a.diag[c], pc, nb) : 0.0f;
            q[c] = qc;
            pq = fmaf(pc, qc, pq);
        }
        block_sum_max(pq, dummy, red, tid);
        const float alpha = (pq != 0.0f) ? __fdividef(rr, pq) : 0.0f;
        float rr_new = 0.0f;
        rmax = 0.0f;
        for (int c = tid; c < N; c += 1024) {
            x[c] = fmaf(alpha, p[c], x[c]);
            const float rc = fmaf(-alpha, q[c], r[c]);
            r[c] = rc;
            rr_new = fmaf(rc, rc, rr_new); rmax = fmaxf(rmax, fabsf(rc));
        }
        block_sum_max(rr_new, rmax, red, tid);        // (its barriers also order the stencil reads above against the p update below)
        const float beta = (rr != 0.0f) ? __fdividef(rr_new, rr) : 0.0f;
        rr = rr_new;
        for (int c = tid; c < N; c += 1024) p[c] = fmaf(beta, p[c], r[c]);
        ++it;
    }
    if (a.iters && tid == 0) a.iters[b] = it;
    __syncthreads();
    if (MODE == 0) {
        float* po = a.p_out + (size_t)b * N;
        for (int c = tid; c < N; c += 1024) po[c] = a.active[c] ? x[c] : -rhs[c] / a.diag[c];
        return;
    }
    float* vyo = a.vy_out + (size_t)b * NY;
    float* vxo = a.vx_out + (size_t)b * NX;
    for (int f = tid; f < (Y + 1) * X; f += 1024) {   // y faces: p = 0 outside the domain
        const int j = f / X, i = f - j * X;
        const float hi = j < Y ? x[j * X + i] : 0.0f, lo = j > 0 ? x[(j - 1) * X + i] : 0.0f;
        vyo[f] = a.my[f] * (vy[f] - (hi - lo));
    }
    for (int f = tid; f < Y * (X + 1); f += 1024) {   // x faces
        const int j = f / (X + 1), i = f - j * (X + 1);
        const float hi = i < X ? x[j * X + i] : 0.0f, lo = i > 0 ? x[j * X + i - 1] : 0.0f;
        vxo[f] = a.mx[f] * (vx[f] - (hi - lo));
    }
    if (a.p_out) {
        float* po = a.p_out + (size_t)b * N;
        for (int c = tid; c < N; c += 1024) po[c] = x[c];
    }
}

static int launch_cg_any(const sol_plan* p, const CgArgs& a, cudaStream_t st, int mode) {
    sol_plan* mp = const_cast<sol_plan*>(p);
    if (!mp->cg_any_scratch) {          // first use (never inside a stream capture: the first call with a set of buffers is eager)
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone)
            return fail(SOL_ERR_UNSUPPORTED, "cg (generic grid): the scratch buffer must be allocated by an eager call first");
        SOL_CUDA(cudaMalloc((void**)&mp->cg_any_scratch, (size_t)p->B_max * 4 * p->Y * p->X * sizeof(float)));
    }
    if (mode == 0) SOL_CUDA(launch_kernel(k_cg_any<0>, dim3(a.B), dim3(1024), 0, st, a, mp->cg_any_scratch));
    else SOL_CUDA(launch_kernel(k_cg_any<1>, dim3(a.B), dim3(1024), 0, st, a, mp->cg_any_scratch));
    SOL_LAUNCHED();
    return SOL_OK;
}

template <int X, int MODE>
static int dispatch_cg(const CgArgs& a, cudaStream_t st, int CL, int R, int TY) {
    // R <= 8: up to 1024 threads (64 registers each); R == 16: up to 512 threads (128 registers)
#define SOL_CG_CASE(RR, CC, NTT) \
    if (R == RR && CL == CC) return launch_cg_t<X, RR, CC, MODE, NTT>(a, st, TY);
    SOL_CG_CASE(2, 1, 1024) SOL_CG_CASE(4, 1, 1024) SOL_CG_CASE(8, 1, 1024) SOL_CG_CASE(16, 1, 512)
    SOL_CG_CASE(2, 2, 1024) SOL_CG_CASE(4, 2, 1024) SOL_CG_CASE(8, 2, 1024) SOL_CG_CASE(16, 2, 512)
    SOL_CG_CASE(2, 4, 1024) SOL_CG_CASE(4, 4, 1024) SOL_CG_CASE(8, 4, 1024) SOL_CG_CASE(16, 4, 512)
    SOL_CG_CASE(2, 8, 1024) SOL_CG_CASE(4, 8, 1024) SOL_CG_CASE(8, 8, 1024) SOL_CG_CASE(16, 8, 512)
#undef SOL_CG_CASE
    return fail(SOL_ERR_UNSUPPORTED, "cg: no kernel for this (rows/thread, cluster) combination");
}

// choose (CL, R, TY) with Y = CL*TY*R; threads = X*TY <= 1024 (R <= 8) or <= 512 (R == 16)
static bool cg_geometry(int Y, int X, int want_cl, int want_r, int& CL, int& R, int& TY) {
    if (!(X == 32 || X == 64 || X == 128)) return false;
    const int cls[4] = {1, 2, 4, 8};
    const int rs[4] = {8, 4, 16, 2};      // preference order
    for (int ci = 0; ci < 4; ++ci) {
        const int cl = cls[ci];
        if (want_cl > 0 && cl != want_cl) continue;
        if (Y % cl) continue;
        const int rows = Y / cl;
        for (int ri = 0; ri < 4; ++ri) {
            const int rr = rs[ri];
            if (want_r > 0 && rr != want_r) continue;
            if (rows % rr) continue;
            const int ty = rows / rr;
            const int max_threads = (rr == 16) ? 512 : 1024;
            if (ty >= 1 && X * ty <= max_threads) { CL = cl; R = rr; TY = ty; return true; }
        }
    }
    return false;
}

// The direct solver is built on first use (host precomputation, a fraction of a second) — never during a stream capture,
// where the iterative path is used instead (the engine's first call with a set of buffers is always eager).
bool direct_active(const sol_plan* p) {
    if (!p->direct_solve || p->boundary != SOL_BOUNDARY_OPEN || p->cluster > 1 || !direct_supported(p)) return false;
    if (!p->dir.tried) {
        sol_plan* mp = const_cast<sol_plan*>(p);
        if (direct_build(mp) != SOL_OK) { direct_free(mp); mp->dir.tried = true; }
    }
    return p->dir.valid;
}

// One cluster per simulation is latency-optimal for a handful of simulations; with one simulation per SM (B = 148) the multigrid
// CG kernel, which keeps a whole solve inside one CTA, has the higher throughput (measured 114 vs 180 us per launch).
// (grids the multigrid kernel does not cover keep the direct solver at every batch size)
bool direct_for_batch(const sol_plan* p, int B) { return (B <= 64 || !mg_supported(p)) && direct_active(p); }

// the solver launch_cg() picks for this batch takes the fused feature I/O: the direct projection, or the compile-time-hierarchy
// multigrid kernel (the same predicate order as launch_cg)
bool cg_fuses(const sol_plan* p, int B) {
    if (direct_for_batch(p, B)) return true;
    return p->boundary == SOL_BOUNDARY_OPEN && p->cg_precond && p->cluster <= 1 && mg_supported(p) && mg3_selected(p);
}

int launch_cg(const sol_plan* p, cudaStream_t st, int B, int mode, const float* rhs, float* p_out, const float* vy,
              const float* vx, float* vy_out, float* vx_out, int* iters, const CgFuse* fuse) {
    if (p->boundary != SOL_BOUNDARY_OPEN) return fail(SOL_ERR_UNSUPPORTED, "pressure solve requires an OPEN-boundary plan");
    bool may_build = true;
    if (!p->dir.tried) {      // the precomputation allocates and copies: never inside a stream capture (the iterative path runs instead)
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) may_build = false;
    }
    if (may_build && direct_for_batch(p, B)) return launch_direct(p, st, B, mode, rhs, p_out, vy, vx, vy_out, vx_out, iters, fuse);
    if (p->cg_precond && p->cluster <= 1 && mg_supported(p))
        return launch_cg_mg(p, st, B, mode, rhs, p_out, vy, vx, vy_out, vx_out, iters, fuse);
    if (fuse && (fuse->feat_out || fuse->gfeat_in)) return fail(SOL_ERR_UNSUPPORTED, "cg: fused feature I/O is not available for this solver variant (see cg_fuses)");
    CgArgs a;
    a.Y = p->Y; a.X = p->X; a.B = B;
    a.diag = p->diag; a.active = p->active; a.my = p->face_my; a.mx = p->face_mx;
    a.rhs = rhs; a.p_out = p_out; a.vy_in = vy; a.vx_in = vx; a.vy_out = vy_out; a.vx_out = vx_out;
    a.tol_abs = p->tol_abs; a.tol_rel = p->tol_rel; a.max_it = p->max_it; a.iters = iters;
    int CL, R, TY;
    if (!cg_geometry(p->Y, p->X, p->cluster, p->cg_rows, CL, R, TY)) {
        if (p->cluster > 1 || p->cg_rows > 0)
            return fail(SOL_ERR_UNSUPPORTED, "cg: an explicit cluster / rows-per-thread setting needs X in {32,64,128} and Y = cluster*TY*R with R in {2,4,8,16}, X*TY <= 1024");
        return launch_cg_any(p, a, st, mode);          // any other grid: the generic one-CTA-per-simulation kernel
    }
#define SOL_CG_X(XX)                                          \
    if (p->X == XX) {                                         \
        if (mode == 0) return dispatch_cg<XX, 0>(a, st, CL, R, TY); \
        return dispatch_cg<XX, 1>(a, st, CL, R, TY);          \
    }
    SOL_CG_X(32) SOL_CG_X(64) SOL_CG_X(128)
#undef SOL_CG_X
    return fail(SOL_ERR_UNSUPPORTED, "cg: unsupported X");
}

}  // namespace sol
