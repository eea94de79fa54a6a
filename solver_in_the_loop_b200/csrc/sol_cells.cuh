// Per-cell arithmetic of the solver stages (forward and adjoint), written once as
// __host__ __device__ functions so that (a) the sm_100a kernels in sol_stencil.cu are thin
// grid-stride loops around them and (b) tests/host_emu can compile the *same* functions with
// g++ and compare them with the CPU oracle when no GPU is present (test infrastructure only:
// the product library never runs these on the host).
//
// Semantics follow SURVEY.md Appendix A (index-space restatement of KarmanFlow.step,
// reference karman-2d/karman_train.py:166-185 + [PHI-RECALL] phiflow 1.5.1):
//   vy [Y+1, X]  y-faces,  vx [Y, X+1]  x-faces,  rho/p/d [Y, X]; row-major, x fastest.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define SOL_HD __host__ __device__ __forceinline__
#else
#define SOL_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define SOL_ATOMIC_ADD(ptr, val) atomicAdd((ptr), (val))
#else
#define SOL_ATOMIC_ADD(ptr, val) (*(ptr) += (val))
#endif

namespace sol {

enum { WRAP_REPLICATE = 0, WRAP_PERIODIC = 1, WRAP_ZERO = 2 };

SOL_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
SOL_HD int modi(int v, int n) { int r = v % n; return r < 0 ? r + n : r; }

// ---------------------------------------------------------------------------------------------
// 5-point Laplace, replicate ("boundary") or periodic padding, unit cells (Appendix A item 1)
// ---------------------------------------------------------------------------------------------
template <int WRAP>
SOL_HD float lap5(const float* __restrict__ c, int H, int W, int j, int i) {
    int ju, jd, il, ir;
    if (WRAP == WRAP_PERIODIC) {
        ju = (j + 1 == H) ? 0 : j + 1; jd = (j == 0) ? H - 1 : j - 1;
        ir = (i + 1 == W) ? 0 : i + 1; il = (i == 0) ? W - 1 : i - 1;
    } else {
        ju = (j + 1 < H) ? j + 1 : H - 1; jd = (j > 0) ? j - 1 : 0;
        ir = (i + 1 < W) ? i + 1 : W - 1; il = (i > 0) ? i - 1 : 0;
    }
    const float cc = c[j * W + i];
    return (c[ju * W + i] + c[jd * W + i] + c[j * W + ir] + c[j * W + il]) - 4.0f * cc;
}

// forward: c + alpha*lap(c), then (optional) Dirichlet BC  c*(1-m)+v   (karman_train.py:175-181)
SOL_HD float diffuse_bc_cell(const float* __restrict__ c, int H, int W, int j, int i, float alpha,
                             const float* __restrict__ bc_mask, const float* __restrict__ bc_val) {
    float out = c[j * W + i] + alpha * lap5<WRAP_REPLICATE>(c, H, W, j, i);
    if (bc_mask) out = out * (1.0f - bc_mask[j * W + i]) + bc_val[j * W + i];
    return out;
}

// adjoint: the replicate-padded Laplace is symmetric (Neumann), so  g_in = (I + alpha*L)((1-m) g)
SOL_HD float diffuse_bc_bwd_cell(const float* __restrict__ g, int H, int W, int j, int i, float alpha,
                                 const float* __restrict__ bc_mask) {
    const int ju = (j + 1 < H) ? j + 1 : H - 1, jd = (j > 0) ? j - 1 : 0;
    const int ir = (i + 1 < W) ? i + 1 : W - 1, il = (i > 0) ? i - 1 : 0;
    float hc = g[j * W + i], hu = g[ju * W + i], hd = g[jd * W + i], hl = g[j * W + il], hr = g[j * W + ir];
    if (bc_mask) {
        hc *= 1.0f - bc_mask[j * W + i];
        hu *= 1.0f - bc_mask[ju * W + i];
        hd *= 1.0f - bc_mask[jd * W + i];
        hl *= 1.0f - bc_mask[j * W + il];
        hr *= 1.0f - bc_mask[j * W + ir];
    }
    return hc + alpha * ((hu + hd + hl + hr) - 4.0f * hc);
}

// ---------------------------------------------------------------------------------------------
// bilinear sampling: weights from the unclamped fraction, corners clamped / wrapped / zeroed
// independently (Appendix A item 3)
// ---------------------------------------------------------------------------------------------
struct Bilerp {
    int j0, j1, i0, i1;
    float wy, wx;
    float m00, m01, m10, m11;   // corner validity (WRAP_ZERO only; 1 otherwise)
};

template <int WRAP>
SOL_HD Bilerp bilerp_setup(float py, float px, int H, int W) {
    // guard against NaN/inf coordinates (diverged velocities) so indices stay in range
    py = fminf(fmaxf(py, -1.0e6f), 1.0e6f);
    px = fminf(fmaxf(px, -1.0e6f), 1.0e6f);
    const float fy = floorf(py), fx = floorf(px);
    Bilerp b;
    b.wy = py - fy; b.wx = px - fx;
    const int j = (int)fy, i = (int)fx;
    b.m00 = b.m01 = b.m10 = b.m11 = 1.0f;
    if (WRAP == WRAP_PERIODIC) {
        b.j0 = modi(j, H); b.j1 = modi(j + 1, H); b.i0 = modi(i, W); b.i1 = modi(i + 1, W);
    } else {
        b.j0 = clampi(j, 0, H - 1); b.j1 = clampi(j + 1, 0, H - 1);
        b.i0 = clampi(i, 0, W - 1); b.i1 = clampi(i + 1, 0, W - 1);
        if (WRAP == WRAP_ZERO) {
            const float y0 = (j >= 0 && j <= H - 1) ? 1.0f : 0.0f, y1 = (j + 1 >= 0 && j + 1 <= H - 1) ? 1.0f : 0.0f;
            const float x0 = (i >= 0 && i <= W - 1) ? 1.0f : 0.0f, x1 = (i + 1 >= 0 && i + 1 <= W - 1) ? 1.0f : 0.0f;
            b.m00 = y0 * x0; b.m01 = y0 * x1; b.m10 = y1 * x0; b.m11 = y1 * x1;
        }
    }
    return b;
}

SOL_HD float bilerp_eval(const float* __restrict__ f, int W, const Bilerp& b) {
    const float v00 = f[b.j0 * W + b.i0] * b.m00, v01 = f[b.j0 * W + b.i1] * b.m01;
    const float v10 = f[b.j1 * W + b.i0] * b.m10, v11 = f[b.j1 * W + b.i1] * b.m11;
    return (1.0f - b.wy) * ((1.0f - b.wx) * v00 + b.wx * v01) + b.wy * ((1.0f - b.wx) * v10 + b.wx * v11);
}

// scatter g * weights into gf (adjoint w.r.t. the sampled field)
SOL_HD void bilerp_scatter(float* gf, int W, const Bilerp& b, float g) {
    SOL_ATOMIC_ADD(&gf[b.j0 * W + b.i0], (1.0f - b.wy) * (1.0f - b.wx) * b.m00 * g);
    SOL_ATOMIC_ADD(&gf[b.j0 * W + b.i1], (1.0f - b.wy) * b.wx * b.m01 * g);
    SOL_ATOMIC_ADD(&gf[b.j1 * W + b.i0], b.wy * (1.0f - b.wx) * b.m10 * g);
    SOL_ATOMIC_ADD(&gf[b.j1 * W + b.i1], b.wy * b.wx * b.m11 * g);
}

// d(sample)/d(py), d(sample)/d(px): derivative of the weights only (floor has zero gradient)
SOL_HD void bilerp_dcoord(const float* __restrict__ f, int W, const Bilerp& b, float& dpy, float& dpx) {
    const float v00 = f[b.j0 * W + b.i0] * b.m00, v01 = f[b.j0 * W + b.i1] * b.m01;
    const float v10 = f[b.j1 * W + b.i0] * b.m10, v11 = f[b.j1 * W + b.i1] * b.m11;
    dpy = (1.0f - b.wx) * (v10 - v00) + b.wx * (v11 - v01);
    dpx = (1.0f - b.wy) * (v01 - v00) + b.wy * (v11 - v10);
}

// ---------------------------------------------------------------------------------------------
// semi-Lagrangian self-advection of the staggered velocity (Appendix A item 3), s = dt/dx
// ---------------------------------------------------------------------------------------------
template <int WRAP>
SOL_HD float advect_vy_cell(const float* __restrict__ vy, const float* __restrict__ vx, int Y, int X, int j, int i, float s) {
    const float uy = vy[j * X + i];
    const Bilerp bu = bilerp_setup<WRAP>((float)j - 0.5f, (float)i + 0.5f, Y, X + 1);
    const float ux = bilerp_eval(vx, X + 1, bu);
    const Bilerp bs = bilerp_setup<WRAP>((float)j - s * uy, (float)i - s * ux, Y + 1, X);
    return bilerp_eval(vy, X, bs);
}

template <int WRAP>
SOL_HD float advect_vx_cell(const float* __restrict__ vy, const float* __restrict__ vx, int Y, int X, int j, int i, float s) {
    const float ux = vx[j * (X + 1) + i];
    const Bilerp bu = bilerp_setup<WRAP>((float)j + 0.5f, (float)i - 0.5f, Y + 1, X);
    const float uy = bilerp_eval(vy, X, bu);
    const Bilerp bs = bilerp_setup<WRAP>((float)j - s * uy, (float)i - s * ux, Y, X + 1);
    return bilerp_eval(vx, X + 1, bs);
}

// density at cell centres; corners outside the array contribute zero (constant extrapolation)
SOL_HD float advect_rho_cell(const float* __restrict__ rho, const float* __restrict__ vy, const float* __restrict__ vx,
                             int Y, int X, int j, int i, float s) {
    const float uy = 0.5f * (vy[j * X + i] + vy[(j + 1) * X + i]);
    const float ux = 0.5f * (vx[j * (X + 1) + i] + vx[j * (X + 1) + i + 1]);
    const Bilerp bs = bilerp_setup<WRAP_ZERO>((float)j - s * uy, (float)i - s * ux, Y, X);
    return bilerp_eval(rho, X, bs);
}

// adjoint of advect_vy_cell for upstream gradient g at y-face (j,i): scatter-adds into gvy, gvx
template <int WRAP>
SOL_HD void advect_vy_cell_bwd(const float* __restrict__ vy, const float* __restrict__ vx, int Y, int X, int j, int i,
                               float s, float g, float* gvy, float* gvx) {
    const float uy = vy[j * X + i];
    const Bilerp bu = bilerp_setup<WRAP>((float)j - 0.5f, (float)i + 0.5f, Y, X + 1);
    const float ux = bilerp_eval(vx, X + 1, bu);
    const Bilerp bs = bilerp_setup<WRAP>((float)j - s * uy, (float)i - s * ux, Y + 1, X);
    bilerp_scatter(gvy, X, bs, g);
    float dpy, dpx;
    bilerp_dcoord(vy, X, bs, dpy, dpx);
    SOL_ATOMIC_ADD(&gvy[j * X + i], -s * dpy * g);
    bilerp_scatter(gvx, X + 1, bu, -s * dpx * g);
}

template <int WRAP>
SOL_HD void advect_vx_cell_bwd(const float* __restrict__ vy, const float* __restrict__ vx, int Y, int X, int j, int i,
                               float s, float g, float* gvy, float* gvx) {
    const float ux = vx[j * (X + 1) + i];
    const Bilerp bu = bilerp_setup<WRAP>((float)j + 0.5f, (float)i - 0.5f, Y + 1, X);
    const float uy = bilerp_eval(vy, X, bu);
    const Bilerp bs = bilerp_setup<WRAP>((float)j - s * uy, (float)i - s * ux, Y, X + 1);
    bilerp_scatter(gvx, X + 1, bs, g);
    float dpy, dpx;
    bilerp_dcoord(vx, X + 1, bs, dpy, dpx);
    SOL_ATOMIC_ADD(&gvx[j * (X + 1) + i], -s * dpx * g);
    bilerp_scatter(gvy, X, bu, -s * dpy * g);
}

// ---------------------------------------------------------------------------------------------
// projection pieces (Appendix A items 5-6) — used by the stand-alone divergence / gradient
// entry points and by the host emulation; the fused CG kernel inlines the same expressions.
// ---------------------------------------------------------------------------------------------
SOL_HD float divergence_cell(const float* __restrict__ vy, const float* __restrict__ vx,
                             const float* __restrict__ my, const float* __restrict__ mx, int Y, int X, int j, int i) {
    const float yl = vy[j * X + i] * my[j * X + i], yh = vy[(j + 1) * X + i] * my[(j + 1) * X + i];
    const float xl = vx[j * (X + 1) + i] * mx[j * (X + 1) + i], xh = vx[j * (X + 1) + i + 1] * mx[j * (X + 1) + i + 1];
    return (yh - yl) + (xh - xl);
}

// vy_out = my * (vy - (p[j,i] - p[j-1,i])), p = 0 outside the domain
SOL_HD float gradsub_vy_cell(const float* __restrict__ vy, const float* __restrict__ p, const float* __restrict__ my,
                             int Y, int X, int j, int i) {
    const float ph = (j < Y) ? p[j * X + i] : 0.0f, pl = (j > 0) ? p[(j - 1) * X + i] : 0.0f;
    return my[j * X + i] * (vy[j * X + i] - (ph - pl));
}

SOL_HD float gradsub_vx_cell(const float* __restrict__ vx, const float* __restrict__ p, const float* __restrict__ mx,
                             int Y, int X, int j, int i) {
    const float ph = (i < X) ? p[j * X + i] : 0.0f, pl = (i > 0) ? p[j * X + i - 1] : 0.0f;
    return mx[j * (X + 1) + i] * (vx[j * (X + 1) + i] - (ph - pl));
}

// y = A p on one cell: active neighbours (solid and outside inactive), diag = #accessible (>=1)
SOL_HD float laplace_cell(const float* __restrict__ p, const unsigned char* __restrict__ active,
                          const float* __restrict__ diag, int Y, int X, int j, int i) {
    const int c = j * X + i;
    if (!active[c]) return -diag[c] * p[c];
    float nb = 0.0f;
    if (j + 1 < Y && active[c + X]) nb += p[c + X];
    if (j > 0 && active[c - X]) nb += p[c - X];
    if (i + 1 < X && active[c + 1]) nb += p[c + 1];
    if (i > 0 && active[c - 1]) nb += p[c - 1];
    return nb - diag[c] * p[c];
}

}  // namespace sol
