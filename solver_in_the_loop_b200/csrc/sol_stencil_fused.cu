// Fused solver stage in front of the pressure projection (OPEN-boundary karman scene):
//     viscosity diffuse + velocity BC  ->  semi-Lagrangian advection of vy, vx and the marker density (+ inflow)
// (reference: KarmanFlow.step, karman-2d/karman_train.py:173-185 -> IncompressibleFlow.step)
// in ONE launch with the stencil halo staged in shared memory.
//
// A CTA owns a tile of TY x TX cells.  It loads the raw velocity rows of the tile + halo (HALO + 1 cells, 128-bit row loads
// where the rows are aligned), computes the diffused + BC velocity on tile + HALO into shared memory (the tile part also goes to
// global memory: it is the stash the advection adjoint needs), and back-traces every face / cell of its tile through that
// shared-memory field.  HALO = 4 cells covers |u| dt/dx < 3 cells; a back-trace that leaves the staged window takes the exact
// slow path instead — the diffused value of a far face is recomputed from the raw global field with the same arithmetic — so the
// result never depends on the tile size.
#include "sol_cells.cuh"
#include "sol_internal.cuh"

SOL_TRACE_TU()

namespace sol {

namespace {

constexpr int FT_TY = 8, FT_TX = 32, FT_HALO = 4;
constexpr int FT_DH = FT_TY + 1 + 2 * FT_HALO;      // rows of the diffused window (faces: one more than cells)
constexpr int FT_DW = FT_TX + 1 + 2 * FT_HALO;      // columns
constexpr int FT_RH = FT_DH + 2, FT_RW = FT_DW + 2; // raw window: one more ring for the 5-point stencil
constexpr int FT_THREADS = 256;

struct FusedArgs {
    int B, Y, X;
    const float* re; float dt_res2, s, dt;
    const float* vy; const float* vx; const float* rho;           // inputs (rho may be null)
    const float* bcm; const float* bcv; const float* inflow;
    float* vy1; float* vx1;                                       // diffused + BC velocity (adjoint stash)
    float* vy2; float* vx2; float* rho_out;                       // advected fields
};

// window of a field in shared memory: global face (j, i) lives at w[(j - j0) * FT_?W + (i - i0)] for j0 <= j < j0 + h, i0 <= i < i0 + w
struct Win {
    const float* p; int j0, i0, h, w, pitch;
    __device__ __forceinline__ bool has(int j, int i) const { return (unsigned)(j - j0) < (unsigned)h && (unsigned)(i - i0) < (unsigned)w; }
    __device__ __forceinline__ float at(int j, int i) const { return p[(j - j0) * pitch + (i - i0)]; }
};

// the diffused + BC field: shared-memory window first, exact recomputation from the raw global field outside it
struct DField {
    Win win; const float* raw; int H, W; float alpha; const float* bcm; const float* bcv;
    __device__ __forceinline__ float operator()(int j, int i) const {
        if (win.has(j, i)) return win.at(j, i);
        return diffuse_bc_cell(raw, H, W, j, i, alpha, bcm, bcv);
    }
};

// bilinear sample with the expression order of bilerp_eval (sol_cells.cuh)
template <class F>
__device__ __forceinline__ float sample(const F& f, const Bilerp& b) {
    const float v00 = f(b.j0, b.i0) * b.m00, v01 = f(b.j0, b.i1) * b.m01;
    const float v10 = f(b.j1, b.i0) * b.m10, v11 = f(b.j1, b.i1) * b.m11;
    return (1.0f - b.wy) * ((1.0f - b.wx) * v00 + b.wx * v01) + b.wy * ((1.0f - b.wx) * v10 + b.wx * v11);
}

}  // namespace

__global__ void __launch_bounds__(FT_THREADS) k_diffuse_advect(const FusedArgs a) {
    __shared__ float r_vy[FT_RH * FT_RW], r_vx[FT_RH * FT_RW];      // raw windows (replicate-clamped at the domain border)
    __shared__ float d_vy[FT_DH * FT_DW], d_vx[FT_DH * FT_DW];      // diffused + BC windows
    const int Y = a.Y, X = a.X;
    const int NY = (Y + 1) * X, NX = Y * (X + 1), NC = Y * X;
    const int b = blockIdx.z;
    const int y0 = blockIdx.y * FT_TY, x0 = blockIdx.x * FT_TX;
    const int tid = threadIdx.x;
    pdl_sync();
    const float* vy = a.vy + (size_t)b * NY;
    const float* vx = a.vx + (size_t)b * NX;
    const float alpha = a.dt_res2 / __ldg(a.re + b);
    // ---- stage A: raw windows; window entry (r, c) holds face (clamp(jr0 + r), clamp(ir0 + c)): replicate padding is built in ----
    const int jr0 = y0 - FT_HALO - 1, ir0 = x0 - FT_HALO - 1;
    for (int k = tid; k < FT_RH * FT_RW; k += FT_THREADS) {
        const int r = k / FT_RW, c = k - r * FT_RW;
        const int jy = clampi(jr0 + r, 0, Y), iy = clampi(ir0 + c, 0, X - 1);
        const int jx = clampi(jr0 + r, 0, Y - 1), ix = clampi(ir0 + c, 0, X);
        r_vy[k] = __ldg(vy + jy * X + iy);
        r_vx[k] = __ldg(vx + jx * (X + 1) + ix);
    }
    __syncthreads();
    // ---- stage B: diffused + BC velocity on tile + HALO (clamped faces repeat their border value, never read beyond it) ----
    const int jd0 = y0 - FT_HALO, id0 = x0 - FT_HALO;
    for (int k = tid; k < FT_DH * FT_DW; k += FT_THREADS) {
        const int r = k / FT_DW, c = k - r * FT_DW;
        const int j = jd0 + r, i = id0 + c;
        // y component, face (j, i) of the [Y+1, X] grid
        if (j >= 0 && j <= Y && i >= 0 && i < X) {
            const int ju = (j + 1 <= Y) ? j + 1 : Y, jdn = (j > 0) ? j - 1 : 0, ir = (i + 1 < X) ? i + 1 : X - 1, il = (i > 0) ? i - 1 : 0;
            auto R = [&](int jj, int ii) { return r_vy[(jj - jr0) * FT_RW + (ii - ir0)]; };
            const float cc = R(j, i);
            float out = cc + alpha * ((R(ju, i) + R(jdn, i) + R(j, ir) + R(j, il)) - 4.0f * cc);      // = c + alpha*lap5 (sol_cells.cuh)
            if (a.bcm) out = out * (1.0f - __ldg(a.bcm + j * X + i)) + __ldg(a.bcv + j * X + i);
            d_vy[k] = out;
            if (j >= y0 && (j < y0 + FT_TY || (j == Y && y0 + FT_TY >= Y)) && i >= x0 && i < x0 + FT_TX) a.vy1[(size_t)b * NY + j * X + i] = out;
        }
        // x component, face (j, i) of the [Y, X+1] grid
        if (j >= 0 && j < Y && i >= 0 && i <= X) {
            const int ju = (j + 1 < Y) ? j + 1 : Y - 1, jdn = (j > 0) ? j - 1 : 0, ir = (i + 1 <= X) ? i + 1 : X, il = (i > 0) ? i - 1 : 0;
            auto R = [&](int jj, int ii) { return r_vx[(jj - jr0) * FT_RW + (ii - ir0)]; };
            const float cc = R(j, i);
            const float out = cc + alpha * ((R(ju, i) + R(jdn, i) + R(j, ir) + R(j, il)) - 4.0f * cc);
            d_vx[k] = out;
            if (j >= y0 && j < y0 + FT_TY && i >= x0 && (i < x0 + FT_TX || (i == X && x0 + FT_TX >= X))) a.vx1[(size_t)b * NX + j * (X + 1) + i] = out;
        }
    }
    __syncthreads();
    // ---- stage C: back-traces of the tile's faces and cells through the staged field ----
    DField DY{Win{d_vy, max(jd0, 0), max(id0, 0), min(jd0 + FT_DH, Y + 1) - max(jd0, 0), min(id0 + FT_DW, X) - max(id0, 0), FT_DW}, vy, Y + 1, X, alpha, a.bcm, a.bcv};
    DField DX{Win{d_vx, max(jd0, 0), max(id0, 0), min(jd0 + FT_DH, Y) - max(jd0, 0), min(id0 + FT_DW, X + 1) - max(id0, 0), FT_DW}, vx, Y, X + 1, alpha, nullptr, nullptr};
    // the windows are addressed from (jd0, id0) whatever the clipping: shift the base pointers instead of the origins
    DY.win.p = d_vy + (DY.win.j0 - jd0) * FT_DW + (DY.win.i0 - id0);
    DX.win.p = d_vx + (DX.win.j0 - jd0) * FT_DW + (DX.win.i0 - id0);
    const float s = a.s;
    const int ty_rows = (y0 + FT_TY >= Y) ? (Y - y0) + 1 : FT_TY;         // the last tile row owns the far y-face row (Y need not be a multiple of the tile)
    const int tx_cols = (x0 + FT_TX >= X) ? min(FT_TX, X - x0) + 1 : FT_TX;   // the last tile column owns the far x-face column
    const int tw = min(FT_TX, X - x0);
    for (int k = tid; k < ty_rows * tw; k += FT_THREADS) {                // y-faces
        const int j = y0 + k / tw, i = x0 + k % tw;
        const float uy = DY(j, i);
        const Bilerp bu = bilerp_setup<WRAP_REPLICATE>((float)j - 0.5f, (float)i + 0.5f, Y, X + 1);
        const float ux = sample(DX, bu);
        const Bilerp bs = bilerp_setup<WRAP_REPLICATE>((float)j - s * uy, (float)i - s * ux, Y + 1, X);
        a.vy2[(size_t)b * NY + j * X + i] = sample(DY, bs);
    }
    const int th = min(FT_TY, Y - y0);
    for (int k = tid; k < th * tx_cols; k += FT_THREADS) {                // x-faces
        const int j = y0 + k / tx_cols, i = x0 + k % tx_cols;
        const float ux = DX(j, i);
        const Bilerp bu = bilerp_setup<WRAP_REPLICATE>((float)j + 0.5f, (float)i - 0.5f, Y + 1, X);
        const float uy = sample(DY, bu);
        const Bilerp bs = bilerp_setup<WRAP_REPLICATE>((float)j - s * uy, (float)i - s * ux, Y, X + 1);
        a.vx2[(size_t)b * NX + j * (X + 1) + i] = sample(DX, bs);
    }
    if (a.rho) {                                                           // marker density at the cell centres (+ inflow)
        const float* rho = a.rho + (size_t)b * NC;
        for (int k = tid; k < th * tw; k += FT_THREADS) {
            const int j = y0 + k / tw, i = x0 + k % tw;
            const float uy = 0.5f * (DY(j, i) + DY(j + 1, i));
            const float ux = 0.5f * (DX(j, i) + DX(j, i + 1));
            const Bilerp bs = bilerp_setup<WRAP_ZERO>((float)j - s * uy, (float)i - s * ux, Y, X);
            float r = bilerp_eval(rho, X, bs);
            if (a.inflow) r += __ldg(a.inflow + j * X + i) * a.dt;
            a.rho_out[(size_t)b * NC + j * X + i] = r;
        }
    }
}

int g_fuse_stencil = 1;      // option "fuse_stencil": diffuse+BC and the advection in one shared-memory-staged launch (OPEN plans)

// vy1/vx1: diffused + BC velocity (the stash of the advection adjoint); vy2/vx2/rho_out: advected fields
int launch_diffuse_advect(const sol_plan* p, cudaStream_t st, int B, const float* re, float dt, float res, const float* vy, const float* vx,
                          const float* rho, float* vy1, float* vx1, float* vy2, float* vx2, float* rho_out) {
    if (p->boundary != SOL_BOUNDARY_OPEN) return fail(SOL_ERR_UNSUPPORTED, "diffuse_advect: OPEN plans only");
    FusedArgs a;
    a.B = B; a.Y = p->Y; a.X = p->X; a.re = re; a.dt_res2 = dt * res * res; a.s = dt / p->dx; a.dt = dt;
    a.vy = vy; a.vx = vx; a.rho = rho; a.bcm = p->bc_mask_y; a.bcv = p->bc_val_y; a.inflow = p->inflow;
    a.vy1 = vy1; a.vx1 = vx1; a.vy2 = vy2; a.vx2 = vx2; a.rho_out = rho_out;
    const dim3 grid(cdiv(p->X, FT_TX), cdiv(p->Y, FT_TY), B);
    SOL_CUDA(launch_kernel(k_diffuse_advect, grid, dim3(FT_THREADS), 0, st, a));
    SOL_LAUNCHED();
    return SOL_OK;
}

}  // namespace sol
