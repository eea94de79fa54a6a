// Direct pressure projection for OPEN-boundary scenes (reference: the Poisson solve inside
// IncompressibleFlow.step -> divergence_free, karman-2d/karman_train.py:167-168,185; the NumPy path of the apply
// scripts solves the same system with a direct sparse solver, karman_apply.py:39).
//
// The scene's operator A is fixed for a plan, so the solve is precomputed instead of iterated:
//   * without the obstacle, A0 is the 5-point Laplacian of a rectangle with p = 0 one cell outside: diagonalised by the
//     type-I discrete sine transform, A0^-1 D = Sy ((Sy D Sx) / lambda) Sx with the symmetric orthogonal matrices Sy, Sx;
//   * the obstacle changes only k rows of A (its solid cells and their fluid neighbours, k = 164 at 128x64):
//     A = A0 + E R^T  ->  A^-1 d = p0 - (W M) (R^T p0),  p0 = A0^-1 d,  W = A0^-1 E (N x k),  M = (I + R^T W)^-1 (k x k)
//     (capacitance-matrix / Woodbury correction); W M is precomputed on the host in double precision, R^T is sparse.
// Two launches per projection: k_direct_solve (one CTA per simulation: divergence -> four small dense products in shared
// memory -> s = R^T p0) and k_direct_apply (all SMs: p = p0 - (W M) s, gradient subtraction, optional feature output).
// No iteration, no stopping rule: the result is the exact solve up to fp32 round-off (4e-7 relative vs the float64
// sparse LU of the oracle), where the reference's CG stops at max|r| < 1e-5.
#include <cooperative_groups.h>
#include <math.h>

#include <vector>

#include "sol_direct_host.h"
#include "sol_internal.cuh"

SOL_TRACE_TU()

namespace sol {

namespace {

struct DirectArgs {
    int B;
    const float* Sy; const float* Sx; const float* ilam;
    const int* rt_col; const float* rt_val;     // [k][5] sparse rows of R^T (column -1 = unused)
    const float* Wt;                            // [kp][N]:  Wt[q][c] = (W M)[c][q]
    int k, kp;
    const float* my; const float* mx; const unsigned char* active; const float* diag;
    const float* rhs;                           // MODE 0
    const float* vy_in; const float* vx_in;     // MODE 1
    float* p0;                                  // [B][N] scratch
    float* zbuf;                                // [B][N] scratch of the streamed kernel (large grids)
    long long* trace;                           // diagnostics: 16 clock64 stamps per CTA of k_direct_solve (null in production)
    float* p_out; float* vy_out; float* vx_out; int* iters;
    // fused feature I/O (see CgFuse)
    float* feat_out; const float* re; float isy, isx, isr; const float* gfeat_in; int cfeat;
};

// asynchronous 16-byte global -> shared copies for the scene constants: all of a thread's copies are in flight at once (register-staged
// loads are interleaved with their stores in groups of four by ptxas: 2.9 us of prologue for 32 KB, 0.9 us this way)
__device__ __forceinline__ void cp_async16(float* dst, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_drain() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }

// acc[rr][cc] += sum_kk At[kk][r0+rr] * Bm[kk][c0+cc]   (At: leading dimension lda, Bm: ldb; TM = 2 or 4 rows per thread)
template <int TM, int K>
__device__ __forceinline__ void tile_gemm(const float* __restrict__ At, int lda, const float* __restrict__ Bm, int ldb, int r0, int c0,
                                          float (&acc)[TM][4]) {
#pragma unroll 8
    for (int kk = 0; kk < K; ++kk) {
        float a[TM];
        if (TM == 4) {
            const float4 v = *reinterpret_cast<const float4*>(At + kk * lda + r0);
            a[0] = v.x; a[1] = v.y; a[2 % TM] = v.z; a[3 % TM] = v.w;
        } else {
            const float2 v = *reinterpret_cast<const float2*>(At + kk * lda + r0);
            a[0] = v.x; a[1] = v.y;
        }
        const float4 b = *reinterpret_cast<const float4*>(Bm + kk * ldb + c0);
#pragma unroll
        for (int rr = 0; rr < TM; ++rr) {
            acc[rr][0] = fmaf(a[rr], b.x, acc[rr][0]); acc[rr][1] = fmaf(a[rr], b.y, acc[rr][1]);
            acc[rr][2] = fmaf(a[rr], b.z, acc[rr][2]); acc[rr][3] = fmaf(a[rr], b.w, acc[rr][3]);
        }
    }
}

template <int TM>
__device__ __forceinline__ void zero_acc(float (&acc)[TM][4]) {
#pragma unroll
    for (int rr = 0; rr < TM; ++rr) acc[rr][0] = acc[rr][1] = acc[rr][2] = acc[rr][3] = 0.0f;
}

// acc tile -> T[c][r] (transposed, leading dimension ld)
template <int TM>
__device__ __forceinline__ void store_transposed(float* __restrict__ T, int ld, int r0, int c0, const float (&acc)[TM][4]) {
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
        if (TM == 4) *reinterpret_cast<float4*>(T + (c0 + cc) * ld + r0) = make_float4(acc[0][cc], acc[1][cc], acc[2 % TM][cc], acc[3 % TM][cc]);
        else *reinterpret_cast<float2*>(T + (c0 + cc) * ld + r0) = make_float2(acc[0][cc], acc[1][cc]);
    }
}

// ---- even / odd symmetry of the sine transform ----
// S[a][n-1-b] = (-1)^a S[a][b] (0-based): an output of even index only sees the mirror SUMS of the contracted vector, one of odd index
// only the mirror DIFFERENCES, so every product needs half the multiplications once the operand has been folded:
//   fold_rows(M): rows kk < K/2 <- M[kk] + M[K-1-kk],  rows K-1-kk <- M[kk] - M[K-1-kk]     (in place)
// Left transform  (Sy M, rows r0 even / r0+1 odd):  acc[0] += At[kk][r0] * M[kk],  acc[1] += At[kk][r0+1] * M[K-1-kk],  kk < K/2
// Right transform (M Sx, columns c0.. even/odd/even/odd): acc[.][0,2] += sums[m] * Sx[m][c],  acc[.][1,3] += diffs[m] * Sx[m][c],  m < K/2
template <int NT>
__device__ __forceinline__ void fold_rows(float* __restrict__ M, int K, int ld, int tid) {
    const int ld4 = ld / 4;
    for (int e = tid; e < (K / 2) * ld4; e += NT) {
        const int kk = e / ld4, c4 = e - kk * ld4;
        float4* pa = reinterpret_cast<float4*>(M + kk * ld) + c4;
        float4* pb = reinterpret_cast<float4*>(M + (K - 1 - kk) * ld) + c4;
        const float4 u = *pa, v = *pb;
        *pa = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
        *pb = make_float4(u.x - v.x, u.y - v.y, u.z - v.z, u.w - v.w);
    }
}

template <int TM, int K>
__device__ __forceinline__ void tile_gemm_sym_left(const float* __restrict__ At, int lda, const float* __restrict__ Mf, int ldb, int r0, int c0,
                                                   float (&acc)[TM][4]) {
#pragma unroll 8
    for (int kk = 0; kk < K / 2; ++kk) {
        float a[TM];
        if (TM == 4) {
            const float4 v = *reinterpret_cast<const float4*>(At + kk * lda + r0);
            a[0] = v.x; a[1] = v.y; a[2 % TM] = v.z; a[3 % TM] = v.w;
        } else {
            const float2 v = *reinterpret_cast<const float2*>(At + kk * lda + r0);
            a[0] = v.x; a[1] = v.y;
        }
        const float4 bs = *reinterpret_cast<const float4*>(Mf + kk * ldb + c0);
        const float4 bd = *reinterpret_cast<const float4*>(Mf + (K - 1 - kk) * ldb + c0);
#pragma unroll
        for (int rr = 0; rr < TM; ++rr) {
            const float4 bm = (rr & 1) ? bd : bs;        // even rows see the mirror sums, odd rows the differences
            acc[rr][0] = fmaf(a[rr], bm.x, acc[rr][0]); acc[rr][1] = fmaf(a[rr], bm.y, acc[rr][1]);
            acc[rr][2] = fmaf(a[rr], bm.z, acc[rr][2]); acc[rr][3] = fmaf(a[rr], bm.w, acc[rr][3]);
        }
    }
}

template <int TM, int K>
__device__ __forceinline__ void tile_gemm_sym_right(const float* __restrict__ Atf, int lda, const float* __restrict__ Bm, int ldb, int r0, int c0,
                                                    float (&acc)[TM][4]) {
#pragma unroll 8
    for (int m = 0; m < K / 2; ++m) {
        float as[TM], ad[TM];
        if (TM == 4) {
            const float4 u = *reinterpret_cast<const float4*>(Atf + m * lda + r0);
            const float4 v = *reinterpret_cast<const float4*>(Atf + (K - 1 - m) * lda + r0);
            as[0] = u.x; as[1] = u.y; as[2 % TM] = u.z; as[3 % TM] = u.w;
            ad[0] = v.x; ad[1] = v.y; ad[2 % TM] = v.z; ad[3 % TM] = v.w;
        } else {
            const float2 u = *reinterpret_cast<const float2*>(Atf + m * lda + r0);
            const float2 v = *reinterpret_cast<const float2*>(Atf + (K - 1 - m) * lda + r0);
            as[0] = u.x; as[1] = u.y; ad[0] = v.x; ad[1] = v.y;
        }
        const float4 b = *reinterpret_cast<const float4*>(Bm + m * ldb + c0);
#pragma unroll
        for (int rr = 0; rr < TM; ++rr) {              // even columns see the mirror sums, odd columns the differences
            acc[rr][0] = fmaf(as[rr], b.x, acc[rr][0]); acc[rr][1] = fmaf(ad[rr], b.y, acc[rr][1]);
            acc[rr][2] = fmaf(as[rr], b.z, acc[rr][2]); acc[rr][3] = fmaf(ad[rr], b.w, acc[rr][3]);
        }
    }
}

// incoming face values (MODE 1), optionally plus the scaled feature gradient of the correction network (fused feat_bwd)
struct FaceIn {
    const float* vy; const float* vx; const float* gf; float isy, isx; int cfeat; int Y, X;
    __device__ __forceinline__ float y(int j, int i) const {
        float v = vy[j * X + i];
        if (gf && j < Y) v = fmaf(gf[((size_t)j * X + i) * cfeat], isy, v);
        return v;
    }
    __device__ __forceinline__ float x(int j, int i) const {
        float v = vx[j * (X + 1) + i];
        if (gf && i < X) v = fmaf(gf[((size_t)j * X + i) * cfeat + 1], isx, v);
        return v;
    }
};

// masked divergence of four consecutive cells (j, i..i+3), i % 4 == 0: 128-bit loads of the y-face rows and their masks, the five
// x-faces as scalars (row pitch X+1); same arithmetic per cell as divergence_cell (sol_cells.cuh) on FaceIn values
template <int X>
__device__ __forceinline__ void divergence4(const FaceIn& in, const float* __restrict__ my, const float* __restrict__ mx,
                                            const unsigned char* __restrict__ active, int j, int i, float (&d)[4]) {
    const float4 yl4 = *reinterpret_cast<const float4*>(in.vy + j * X + i), yh4 = *reinterpret_cast<const float4*>(in.vy + (j + 1) * X + i);
    const float4 ml4 = __ldg(reinterpret_cast<const float4*>(my + j * X + i)), mh4 = __ldg(reinterpret_cast<const float4*>(my + (j + 1) * X + i));
    float yl[4] = {yl4.x, yl4.y, yl4.z, yl4.w}, yh[4] = {yh4.x, yh4.y, yh4.z, yh4.w};
    const float ml[4] = {ml4.x, ml4.y, ml4.z, ml4.w}, mh[4] = {mh4.x, mh4.y, mh4.z, mh4.w};
    float xv[5], mv[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) { xv[c] = in.vx[j * (X + 1) + i + c]; mv[c] = __ldg(mx + j * (X + 1) + i + c); }
    if (in.gf) {      // fused feat_bwd: the incoming gradient plus the scaled feature gradient (FaceIn::y / ::x)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            yl[c] = fmaf(in.gf[((size_t)j * X + i + c) * in.cfeat], in.isy, yl[c]);
            if (j + 1 < in.Y) yh[c] = fmaf(in.gf[((size_t)(j + 1) * X + i + c) * in.cfeat], in.isy, yh[c]);
            xv[c] = fmaf(in.gf[((size_t)j * X + i + c) * in.cfeat + 1], in.isx, xv[c]);
        }
        if (i + 4 < X) xv[4] = fmaf(in.gf[((size_t)j * X + i + 4) * in.cfeat + 1], in.isx, xv[4]);
    }
    const uchar4 ac = *reinterpret_cast<const uchar4*>(active + j * X + i);
    const unsigned char av[4] = {ac.x, ac.y, ac.z, ac.w};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float v = (mh[c] * yh[c] - ml[c] * yl[c]) + (mv[c + 1] * xv[c + 1] - mv[c] * xv[c]);
        d[c] = av[c] ? v : 0.0f;
    }
}

// A thread-block cluster of CL CTAs per simulation, split along y: p0 = Sy ((Sy D Sx) * ilam) Sx.  Every CTA builds the whole
// right-hand side D (cheap), then owns YS = Y/CL rows of the three products U = Sy D, V = (U Sx) * ilam, Z = V Sx — no exchange
// needed, the row split survives right-multiplications — and pushes its rows of Z into every CTA's copy through distributed
// shared memory for the last product p0 = Sy Z.
#define DSTAMP(slot) do { if (a.trace && threadIdx.x == 0) a.trace[(blockIdx.y * gridDim.x + blockIdx.x) * 16 + (slot)] = clock64(); } while (0)

template <int Y, int X, int CL, int TM, int MODE>
__global__ void __launch_bounds__((Y / CL / TM) * (X / 4), 1) k_direct_solve(const DirectArgs a) {
    namespace cg = cooperative_groups;
    DSTAMP(0);
    constexpr int YS = Y / CL;
    constexpr int NT = (YS / TM) * (X / 4);
    constexpr int N = Y * X;
    extern __shared__ __align__(16) float dsm[];
    float* sSy = dsm;                  // [Y][YS]: sSy[kk][lr] = Sy[kk][rbase + lr]  (Sy is symmetric)
    float* sSx = sSy + Y * YS;         // [X][X]
    float* sD = sSx + X * X;           // [Y][X]  right-hand side
    float* sZ = sD + N;                // [Y][X]  gathered Z
    float* t0 = sZ + N;                // [X][YS] transposed slices
    float* t1 = t0 + X * YS;           // [X][YS]
    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    int rank = 0;
    if (CL > 1) rank = (int)cg::this_cluster().block_rank();
    const int rbase = rank * YS;
    // scene constants first: they do not depend on the previous kernel (programmatic dependent launch)
    for (int k = tid * 4; k < Y * YS; k += NT * 4) {
        const int kk = k / YS, lr = k - kk * YS;
        cp_async16(sSy + k, a.Sy + kk * Y + rbase + lr);
    }
    for (int k = tid * 4; k < X * X; k += NT * 4) cp_async16(sSx + k, a.Sx + k);
    DSTAMP(1);
    pdl_sync();
    cp_async_drain();       // this thread's copies have landed; the barrier below publishes them
    DSTAMP(2);
    if (CL > 1) cg::this_cluster().sync();      // every CTA of the cluster is running before any remote shared-memory access
    DSTAMP(3);
    // ---- right-hand side D[j][i]: every CTA computes its YS rows and pushes them into all copies ----
    {
        FaceIn in{a.vy_in + (size_t)b * (Y + 1) * X, a.vx_in + (size_t)b * Y * (X + 1),
                  a.gfeat_in ? a.gfeat_in + (size_t)b * N * a.cfeat : nullptr, a.isy, a.isx, a.cfeat, Y, X};
        const float* rhs = (MODE == 0) ? a.rhs + (size_t)b * N : nullptr;
        float* dst[CL];
#pragma unroll
        for (int peer = 0; peer < CL; ++peer) dst[peer] = (CL > 1) ? cg::this_cluster().map_shared_rank(sD, peer) : sD;
#pragma unroll 2
        for (int lc = tid * 4; lc < YS * X; lc += NT * 4) {
            const int c = rbase * X + lc;
            const int j = c / X, i = c - j * X;
            float d[4];
            if (MODE == 1) {
                divergence4<X>(in, a.my, a.mx, a.active, j, i, d);
            } else {
                const float4 r4 = *reinterpret_cast<const float4*>(rhs + c);
                const uchar4 ac = *reinterpret_cast<const uchar4*>(a.active + c);
                d[0] = ac.x ? r4.x : 0.0f; d[1] = ac.y ? r4.y : 0.0f; d[2] = ac.z ? r4.z : 0.0f; d[3] = ac.w ? r4.w : 0.0f;
            }
#pragma unroll
            for (int peer = 0; peer < CL; ++peer) *reinterpret_cast<float4*>(dst[peer] + c) = make_float4(d[0], d[1], d[2], d[3]);
        }
    }
    DSTAMP(4);
    if (CL > 1) cg::this_cluster().sync();
    else __syncthreads();
    DSTAMP(5);
    static_assert(TM % 2 == 0 && (Y / CL) % TM == 0 && X % 8 == 0, "the folded products pair even with odd rows / columns");
    const int tx = tid % (X / 4), ty = tid / (X / 4);
    const int r0 = ty * TM, c0 = tx * 4;        // r0: row inside this CTA's slice (even); rbase is even too
    float acc[TM][4];
    // stage 1: U = Sy D (my rows) on the folded D, transposed into t0
    fold_rows<NT>(sD, Y, X, tid);
    __syncthreads();
    zero_acc<TM>(acc);
    tile_gemm_sym_left<TM, Y>(sSy, YS, sD, X, r0, c0, acc);
    store_transposed<TM>(t0, YS, r0, c0, acc);
    __syncthreads();
    fold_rows<NT>(t0, X, YS, tid);
    __syncthreads();
    DSTAMP(6);
    // stage 2: V = (U Sx) * ilam, transposed into t1
    zero_acc<TM>(acc);
    tile_gemm_sym_right<TM, X>(t0, YS, sSx, X, r0, c0, acc);
#pragma unroll
    for (int rr = 0; rr < TM; ++rr) {
        const float4 l = __ldg(reinterpret_cast<const float4*>(a.ilam + (rbase + r0 + rr) * X + c0));
        acc[rr][0] *= l.x; acc[rr][1] *= l.y; acc[rr][2] *= l.z; acc[rr][3] *= l.w;
    }
    store_transposed<TM>(t1, YS, r0, c0, acc);
    __syncthreads();
    fold_rows<NT>(t1, X, YS, tid);
    __syncthreads();
    DSTAMP(7);
    // stage 3: Z = V Sx; my rows go into every CTA's sZ
    zero_acc<TM>(acc);
    tile_gemm_sym_right<TM, X>(t1, YS, sSx, X, r0, c0, acc);
    DSTAMP(8);
    if (CL > 1) {
        cg::cluster_group cluster = cg::this_cluster();
#pragma unroll
        for (int peer = 0; peer < CL; ++peer) {
            float* rz = cluster.map_shared_rank(sZ, peer);
#pragma unroll
            for (int rr = 0; rr < TM; ++rr)
                *reinterpret_cast<float4*>(rz + (rbase + r0 + rr) * X + c0) = make_float4(acc[rr][0], acc[rr][1], acc[rr][2], acc[rr][3]);
        }
        cluster.sync();
    } else {
#pragma unroll
        for (int rr = 0; rr < TM; ++rr) *reinterpret_cast<float4*>(sZ + (r0 + rr) * X + c0) = make_float4(acc[rr][0], acc[rr][1], acc[rr][2], acc[rr][3]);
        __syncthreads();
    }
    DSTAMP(9);
    // stage 4: p0 = Sy Z (my rows) on the folded Z -> global memory
    fold_rows<NT>(sZ, Y, X, tid);
    __syncthreads();
    zero_acc<TM>(acc);
    tile_gemm_sym_left<TM, Y>(sSy, YS, sZ, X, r0, c0, acc);
    float* p0g = a.p0 + (size_t)b * N;
#pragma unroll
    for (int rr = 0; rr < TM; ++rr)
        *reinterpret_cast<float4*>(p0g + (rbase + r0 + rr) * X + c0) = make_float4(acc[rr][0], acc[rr][1], acc[rr][2], acc[rr][3]);
    if (a.iters && tid == 0 && rank == 0) a.iters[b] = 0;      // no iterations: a direct solve
    DSTAMP(10);
}

// Larger grids (256x128): the two operands every CTA needs COMPLETELY — D for U = Sy D and Z for p0 = Sy Z, 128 KB each — do not fit
// next to the transform matrices.  Same four products and the same row split over the cluster, but D and Z live in global
// memory (L2-resident scratch: every CTA writes its rows, cluster barrier) and are streamed through a double-buffered
// shared-memory chunk of KC rows (next chunk in registers while the current one feeds the FMAs: one barrier per chunk).
template <int Y, int X, int CL, int TM, int KC, int MODE>
__global__ void __launch_bounds__((Y / CL / TM) * (X / 4), 1) k_direct_solve_big(const DirectArgs a) {
    namespace cg = cooperative_groups;
    constexpr int YS = Y / CL;
    constexpr int NT = (YS / TM) * (X / 4);
    constexpr int N = Y * X;
    constexpr int CH4 = KC * X / 4 / NT;            // float4 per thread and chunk
    static_assert(KC * X / 4 % NT == 0 && (Y / 2) % KC == 0, "chunk geometry");
    extern __shared__ __align__(16) float dsm[];
    float* sSy = dsm;                  // [Y][YS]: sSy[kk][lr] = Sy[kk][rbase + lr]  (Sy is symmetric)
    float* sSx = sSy + Y * YS;         // [X][X]
    float* t0 = sSx + X * X;           // [X][YS] transposed slices
    float* t1 = t0 + X * YS;           // [X][YS]
    float* ch = t1 + X * YS;           // [2 buffers][sums | differences][KC][X] streamed, folded operand rows
    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int rbase = rank * YS;
    for (int k = tid * 4; k < Y * YS; k += NT * 4) {
        const int kk = k / YS, lr = k - kk * YS;
        cp_async16(sSy + k, a.Sy + kk * Y + rbase + lr);
    }
    for (int k = tid * 4; k < X * X; k += NT * 4) cp_async16(sSx + k, a.Sx + k);
    pdl_sync();
    cp_async_drain();       // published by the cluster barrier behind the right-hand side
    float* Dg = a.p0 + (size_t)b * N;       // the obstacle-free solution's buffer doubles as the scratch of D (dead before p0 is written)
    float* Zg = a.zbuf + (size_t)b * N;
    // ---- right-hand side: my YS rows -> global scratch ----
    {
        FaceIn in{a.vy_in + (size_t)b * (Y + 1) * X, a.vx_in + (size_t)b * Y * (X + 1),
                  a.gfeat_in ? a.gfeat_in + (size_t)b * N * a.cfeat : nullptr, a.isy, a.isx, a.cfeat, Y, X};
        const float* rhs = (MODE == 0) ? a.rhs + (size_t)b * N : nullptr;
#pragma unroll 2
        for (int lc = tid * 4; lc < YS * X; lc += NT * 4) {
            const int c = rbase * X + lc;
            const int j = c / X, i = c - j * X;
            float d[4];
            if (MODE == 1) {
                divergence4<X>(in, a.my, a.mx, a.active, j, i, d);
            } else {
                const float4 r4 = *reinterpret_cast<const float4*>(rhs + c);
                const uchar4 ac = *reinterpret_cast<const uchar4*>(a.active + c);
                d[0] = ac.x ? r4.x : 0.0f; d[1] = ac.y ? r4.y : 0.0f; d[2] = ac.z ? r4.z : 0.0f; d[3] = ac.w ? r4.w : 0.0f;
            }
            *reinterpret_cast<float4*>(Dg + c) = make_float4(d[0], d[1], d[2], d[3]);
        }
    }
    cluster.sync();       // release / acquire at cluster scope: the peers' rows of D are visible
    const int tx = tid % (X / 4), ty = tid / (X / 4);
    const int r0 = ty * TM, c0 = tx * 4;
    float acc[TM][4];
    static_assert(TM % 2 == 0 && (Y / CL) % TM == 0 && X % 8 == 0, "the folded products pair even with odd rows / columns");
    static_assert(TM == 2, "the streamed product is written for row pairs");
    // acc (+)= Sy[my rows, :] * G for a [Y][X] operand G streamed from global memory, with the even / odd fold of the sine transform
    // (see fold_rows): chunk c brings rows kk in [c*KC, (c+1)*KC) AND their mirrors Y-1-kk, folded on the fly into sums / differences
    auto stream_gemm = [&](const float* __restrict__ G) {
        float4 ns[CH4], nd[CH4];
        auto fetch = [&](int c) {
#pragma unroll
            for (int e = 0; e < CH4; ++e) {
                const int idx = tid + e * NT, l = idx / (X / 4), c4 = idx - l * (X / 4);
                const float4 u = __ldcg(reinterpret_cast<const float4*>(G + (size_t)(c * KC + l) * X) + c4);
                const float4 v = __ldcg(reinterpret_cast<const float4*>(G + (size_t)(Y - 1 - c * KC - l) * X) + c4);
                ns[e] = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
                nd[e] = make_float4(u.x - v.x, u.y - v.y, u.z - v.z, u.w - v.w);
            }
        };
        auto stash = [&](int buf) {
#pragma unroll
            for (int e = 0; e < CH4; ++e) {
                reinterpret_cast<float4*>(ch + buf * 2 * KC * X)[tid + e * NT] = ns[e];
                reinterpret_cast<float4*>(ch + buf * 2 * KC * X + KC * X)[tid + e * NT] = nd[e];
            }
        };
        fetch(0);
        stash(0);
        __syncthreads();
#pragma unroll 1
        for (int c = 0; c < Y / 2 / KC; ++c) {
            if (c + 1 < Y / 2 / KC) fetch(c + 1);
            const float* At = sSy + c * KC * YS;
            const float* Ms = ch + (c & 1) * 2 * KC * X;
            const float* Md = Ms + KC * X;
#pragma unroll 8
            for (int l = 0; l < KC; ++l) {
                const float2 av = *reinterpret_cast<const float2*>(At + l * YS + r0);
                const float4 bs = *reinterpret_cast<const float4*>(Ms + l * X + c0);
                const float4 bd = *reinterpret_cast<const float4*>(Md + l * X + c0);
                acc[0][0] = fmaf(av.x, bs.x, acc[0][0]); acc[0][1] = fmaf(av.x, bs.y, acc[0][1]);
                acc[0][2] = fmaf(av.x, bs.z, acc[0][2]); acc[0][3] = fmaf(av.x, bs.w, acc[0][3]);
                acc[1][0] = fmaf(av.y, bd.x, acc[1][0]); acc[1][1] = fmaf(av.y, bd.y, acc[1][1]);
                acc[1][2] = fmaf(av.y, bd.z, acc[1][2]); acc[1][3] = fmaf(av.y, bd.w, acc[1][3]);
            }
            if (c + 1 < Y / 2 / KC) stash((c + 1) & 1);
            __syncthreads();
        }
    };
    // stage 1: U = Sy D (my rows), transposed into t0
    zero_acc<TM>(acc);
    stream_gemm(Dg);
    store_transposed<TM>(t0, YS, r0, c0, acc);
    __syncthreads();
    fold_rows<NT>(t0, X, YS, tid);
    __syncthreads();
    // stage 2: V = (U Sx) * ilam, transposed into t1
    zero_acc<TM>(acc);
    tile_gemm_sym_right<TM, X>(t0, YS, sSx, X, r0, c0, acc);
#pragma unroll
    for (int rr = 0; rr < TM; ++rr) {
        const float4 l = __ldg(reinterpret_cast<const float4*>(a.ilam + (rbase + r0 + rr) * X + c0));
        acc[rr][0] *= l.x; acc[rr][1] *= l.y; acc[rr][2] *= l.z; acc[rr][3] *= l.w;
    }
    store_transposed<TM>(t1, YS, r0, c0, acc);
    __syncthreads();
    fold_rows<NT>(t1, X, YS, tid);
    __syncthreads();
    // stage 3: Z = V Sx; my rows -> global scratch
    zero_acc<TM>(acc);
    tile_gemm_sym_right<TM, X>(t1, YS, sSx, X, r0, c0, acc);
#pragma unroll
    for (int rr = 0; rr < TM; ++rr)
        *reinterpret_cast<float4*>(Zg + (rbase + r0 + rr) * X + c0) = make_float4(acc[rr][0], acc[rr][1], acc[rr][2], acc[rr][3]);
    cluster.sync();       // every CTA is done reading D (p0 may be overwritten) and all of Z is visible
    // stage 4: p0 = Sy Z (my rows) -> global memory
    zero_acc<TM>(acc);
    stream_gemm(Zg);
    float* p0g = a.p0 + (size_t)b * N;
#pragma unroll
    for (int rr = 0; rr < TM; ++rr)
        *reinterpret_cast<float4*>(p0g + (rbase + r0 + rr) * X + c0) = make_float4(acc[rr][0], acc[rr][1], acc[rr][2], acc[rr][3]);
    if (a.iters && tid == 0 && rank == 0) a.iters[b] = 0;      // no iterations: a direct solve
}

// p = p0 - (W M) s on a tile of R rows (+ the row below it), then the gradient subtraction / optional outputs.  The correction sum
// over the changed rows is split over QG thread groups (same cells, disjoint q ranges) and reduced through shared memory.
template <int X, int R, int MODE>
__global__ void __launch_bounds__(4 * (R + 1) * X) k_direct_apply(const DirectArgs a, int Y) {
    constexpr int NCELL = (R + 1) * X, QG = 4;
    __shared__ float sp[QG][NCELL];
    extern __shared__ __align__(16) float st[];      // [kp]
    const int N = Y * X;
    const int b = blockIdx.y;
    const int j0 = blockIdx.x * R;
    const int tid = threadIdx.x;
    pdl_sync();
    // s = R^T p0: the changed rows of the operator applied to the obstacle-free solution (<= 5 entries per row)
    for (int q = tid; q < a.kp; q += QG * NCELL) {
        float sv = 0.0f;
        if (q < a.k) {
#pragma unroll
            for (int e = 0; e < 5; ++e) {
                const int col = __ldg(a.rt_col + q * 5 + e);
                if (col >= 0) sv = fmaf(__ldg(a.rt_val + q * 5 + e), a.p0[(size_t)b * N + col], sv);
            }
        }
        st[q] = sv;
    }
    __syncthreads();
    const int g = tid / NCELL, cell = tid - g * NCELL;
    const int lr = cell / X, i = cell - lr * X;
    const int j = j0 - 1 + lr;                       // local row 0 is the halo row below the tile
    {
        float part = 0.0f;
        if (j >= 0 && j < Y) {
            // kp is a multiple of 32 and the rows q >= k of Wt are zero: 8 independent loads in flight per thread
            const int kq = a.kp / QG;
            const float* w = a.Wt + (size_t)(g * kq) * N + j * X + i;
            const float* sq = st + g * kq;
            float corr[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 1
            for (int q = 0; q < kq; q += 8) {
                float wv[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) wv[e] = __ldg(w + (size_t)(q + e) * N);
#pragma unroll
                for (int e = 0; e < 8; ++e) corr[e & 3] = fmaf(wv[e], sq[q + e], corr[e & 3]);
            }
            part = (corr[0] + corr[1]) + (corr[2] + corr[3]);
        }
        sp[g][cell] = part;
    }
    __syncthreads();
    if (g != 0) return;
    float p = 0.0f;
    if (j >= 0 && j < Y) {
        const int c = j * X + i;
        p = a.p0[(size_t)b * N + c] - ((sp[0][cell] + sp[1][cell]) + (sp[2][cell] + sp[3][cell]));
        if (!a.active[c]) p = (MODE == 0) ? -a.rhs[(size_t)b * N + c] / a.diag[c] : 0.0f;
    }
    sp[0][cell] = p;
    // only the first NCELL threads (whole warps) are left: a named barrier among them
    asm volatile("bar.sync 1, %0;" ::"n"(NCELL) : "memory");
    if (lr == 0 || j >= Y) return;
    const int c = j * X + i;
    if (MODE == 0) { a.p_out[(size_t)b * N + c] = p; return; }
    FaceIn in{a.vy_in + (size_t)b * (Y + 1) * X, a.vx_in + (size_t)b * Y * (X + 1),
              a.gfeat_in ? a.gfeat_in + (size_t)b * N * a.cfeat : nullptr, a.isy, a.isx, a.cfeat, Y, X};
    float* vyo = a.vy_out + (size_t)b * (Y + 1) * X;
    float* vxo = a.vx_out + (size_t)b * Y * (X + 1);
    const float pdn = sp[0][cell - X];                              // p[j-1][i] (0 below the domain)
    const float plf = (i > 0) ? sp[0][cell - 1] : 0.0f;             // p[j][i-1] (0 left of the domain)
    const float oy = a.my[j * X + i] * (in.y(j, i) - (p - pdn));
    const float ox = a.mx[j * (X + 1) + i] * (in.x(j, i) - (p - plf));
    vyo[j * X + i] = oy;
    vxo[j * (X + 1) + i] = ox;
    if (i == X - 1) vxo[j * (X + 1) + X] = a.mx[j * (X + 1) + X] * (in.x(j, X) + p);
    if (j == Y - 1) vyo[Y * X + i] = a.my[Y * X + i] * (in.y(Y, i) + p);
    if (a.p_out) a.p_out[(size_t)b * N + c] = p;
    if (a.feat_out) {
        float* f = a.feat_out + ((size_t)b * N + c) * a.cfeat;
        f[0] = oy * a.isy; f[1] = ox * a.isx; f[2] = a.re[b] * a.isr;
    }
}

template <typename T>
int up(T** dst, const std::vector<T>& src) {
    SOL_CUDA(cudaMalloc((void**)dst, src.size() * sizeof(T)));
    SOL_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return SOL_OK;
}

}  // namespace

long long* g_direct_trace = nullptr;      // diagnostics (sol_debug_direct_trace)

bool direct_supported(const sol_plan* p) {
    return p->boundary == SOL_BOUNDARY_OPEN && ((p->Y == 128 && p->X == 64) || (p->Y == 64 && p->X == 32) || (p->Y == 256 && p->X == 128));
}

void direct_free(sol_plan* p) {
    sol_direct& d = p->dir;
    cudaFree(d.Sy); cudaFree(d.Sx); cudaFree(d.ilam); cudaFree(d.rt_col); cudaFree(d.rt_val); cudaFree(d.Wt);
    cudaFree(d.p0); cudaFree(d.zbuf);
    d = sol_direct();
}

// Host precomputation (sol_direct_host.h, double precision) of the transform matrices and the capacitance correction, then upload.
int direct_build(sol_plan* p) {
    sol_direct& d = p->dir;
    d.tried = true;
    if (!direct_supported(p)) return SOL_OK;
    const int N = p->Y * p->X;
    if ((int)p->h_active.size() != N || (int)p->h_diag.size() != N) return SOL_OK;
    DirectHost h;
    if (!direct_precompute(p->Y, p->X, p->h_active.data(), p->h_diag.data(), h)) return SOL_OK;      // singular: stays disabled
    SOL_TRY(up(&d.Sy, h.Sy)); SOL_TRY(up(&d.Sx, h.Sx)); SOL_TRY(up(&d.ilam, h.ilam));
    SOL_TRY(up(&d.rt_col, h.rt_col)); SOL_TRY(up(&d.rt_val, h.rt_val)); SOL_TRY(up(&d.Wt, h.Wt));
    SOL_CUDA(cudaMalloc((void**)&d.p0, (size_t)p->B_max * N * sizeof(float)));
    if (p->Y * p->X > 128 * 64) SOL_CUDA(cudaMalloc((void**)&d.zbuf, (size_t)p->B_max * N * sizeof(float)));
    d.k = h.k; d.kp = h.kp;
    d.valid = true;
    return SOL_OK;
}

template <int Y, int X, int CL, int TM, int R>
static int launch_direct_t(const DirectArgs& a, cudaStream_t st, int mode) {
    constexpr int YS = Y / CL;
    constexpr int NT = (YS / TM) * (X / 4);
    const size_t smem = (size_t)(Y * YS + X * X + 2 * Y * X + 2 * X * YS) * sizeof(float);
    auto k0 = k_direct_solve<Y, X, CL, TM, 0>;
    auto k1 = k_direct_solve<Y, X, CL, TM, 1>;
    static bool attr_done = false;
    if (!attr_done) {
        SOL_CUDA(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        SOL_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done = true;
    }
    if (smem > 200 * 1024) return fail(SOL_ERR_UNSUPPORTED, "direct solve: scene does not fit the shared-memory kernel");
    {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(CL, a.B); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute attr[2];
        int na = 0;
        if (CL > 1) {
            attr[na].id = cudaLaunchAttributeClusterDimension;
            attr[na].val.clusterDim.x = CL; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
            ++na;
        }
        if (g_pdl) {
            attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[na].val.programmaticStreamSerializationAllowed = 1;
            ++na;
        }
        cfg.attrs = attr; cfg.numAttrs = na;
        if (g_trace_names) trace_record_launch(mode == 0 ? (const void*)k0 : (const void*)k1);
        if (mode == 0) SOL_CUDA(cudaLaunchKernelEx(&cfg, k0, a));
        else SOL_CUDA(cudaLaunchKernelEx(&cfg, k1, a));
        SOL_LAUNCHED();
    }
    const dim3 grid(cdiv(Y, R), a.B), block(4 * (R + 1) * X);
    const size_t smem2 = (size_t)a.kp * sizeof(float);
    if (mode == 0) SOL_CUDA(launch_kernel(k_direct_apply<X, R, 0>, grid, block, smem2, st, a, Y));
    else SOL_CUDA(launch_kernel(k_direct_apply<X, R, 1>, grid, block, smem2, st, a, Y));
    SOL_LAUNCHED();
    return SOL_OK;
}

template <int Y, int X, int CL, int TM, int KC, int R>
static int launch_direct_big_t(const DirectArgs& a, cudaStream_t st, int mode) {
    constexpr int YS = Y / CL;
    constexpr int NT = (YS / TM) * (X / 4);
    const size_t smem = (size_t)(Y * YS + X * X + 2 * X * YS + 4 * KC * X) * sizeof(float);
    auto k0 = k_direct_solve_big<Y, X, CL, TM, KC, 0>;
    auto k1 = k_direct_solve_big<Y, X, CL, TM, KC, 1>;
    SOL_CUDA(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SOL_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(CL, a.B); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute attr[2];
        int na = 0;
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = CL; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        ++na;
        if (g_pdl) {
            attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[na].val.programmaticStreamSerializationAllowed = 1;
            ++na;
        }
        cfg.attrs = attr; cfg.numAttrs = na;
        if (g_trace_names) trace_record_launch(mode == 0 ? (const void*)k0 : (const void*)k1);
        if (mode == 0) SOL_CUDA(cudaLaunchKernelEx(&cfg, k0, a));
        else SOL_CUDA(cudaLaunchKernelEx(&cfg, k1, a));
        SOL_LAUNCHED();
    }
    const dim3 grid(cdiv(Y, R), a.B), block(4 * (R + 1) * X);
    const size_t smem2 = (size_t)a.kp * sizeof(float);
    if (mode == 0) SOL_CUDA(launch_kernel(k_direct_apply<X, R, 0>, grid, block, smem2, st, a, Y));
    else SOL_CUDA(launch_kernel(k_direct_apply<X, R, 1>, grid, block, smem2, st, a, Y));
    SOL_LAUNCHED();
    return SOL_OK;
}

int launch_direct(const sol_plan* p, cudaStream_t st, int B, int mode, const float* rhs, float* p_out, const float* vy, const float* vx,
                  float* vy_out, float* vx_out, int* iters, const CgFuse* fuse) {
    const sol_direct& d = p->dir;
    if (!d.valid) return fail(SOL_ERR_UNSUPPORTED, "direct solve: not available for this plan");
    DirectArgs a;
    memset(&a, 0, sizeof(a));
    a.B = B; a.Sy = d.Sy; a.Sx = d.Sx; a.ilam = d.ilam; a.rt_col = d.rt_col; a.rt_val = d.rt_val; a.Wt = d.Wt; a.k = d.k; a.kp = d.kp;
    a.my = p->face_my; a.mx = p->face_mx; a.active = p->active; a.diag = p->diag;
    a.rhs = rhs; a.vy_in = vy; a.vx_in = vx; a.p0 = d.p0; a.zbuf = d.zbuf; a.p_out = p_out; a.vy_out = vy_out; a.vx_out = vx_out; a.iters = iters;
    a.cfeat = 3;
    if (fuse && (fuse->feat_out || fuse->gfeat_in)) {
        if (mode != 1) return fail(SOL_ERR_UNSUPPORTED, "direct solve: fused feature I/O needs the projection mode");
        if (fuse->cfeat < 3 && fuse->feat_out) return fail(SOL_ERR_UNSUPPORTED, "direct solve: fused feature output needs >= 3 feature channels");
        a.feat_out = fuse->feat_out; a.re = fuse->re; a.isy = fuse->isy; a.isx = fuse->isx; a.isr = fuse->isr;
        a.gfeat_in = fuse->gfeat_in; a.cfeat = fuse->cfeat;
    }
    a.trace = g_direct_trace;
    if (mode == 0 && (!rhs || !p_out)) return fail(SOL_ERR_INVALID, "direct solve: rhs / p_out required");
    if (mode == 1 && (!vy || !vx || !vy_out || !vx_out)) return fail(SOL_ERR_INVALID, "direct solve: velocity pointers required");
    if (p->Y == 128 && p->X == 64) return launch_direct_t<128, 64, 4, 4, 3>(a, st, mode);      // 4x4 register tiles: the folded left products are shared-memory-bandwidth bound
    if (p->Y == 64 && p->X == 32) return launch_direct_t<64, 32, 2, 2, 7>(a, st, mode);
    if (p->Y == 256 && p->X == 128) return launch_direct_big_t<256, 128, 8, 2, 16, 1>(a, st, mode);
    return fail(SOL_ERR_UNSUPPORTED, "direct solve: unsupported grid");
}

}  // namespace sol

// Diagnostics hook (not part of the public ABI): 16 clock64 stamps per CTA of the next k_direct_solve launches; null = off.
extern "C" void sol_debug_direct_trace(long long* buf) { sol::g_direct_trace = buf; }
