// Stencil / elementwise stages of the solver step and their adjoints (sm_100a).
// Every kernel is a coalesced grid-stride loop over the struct-of-arrays face / cell arrays
// (x fastest) around the per-cell functions in sol_cells.cuh.  At the reference grid sizes all
// fields are L2-resident (a [3,128,64] field is 98 KB), so the stencils read neighbours straight
// through L1/L2; HBM traffic is the compulsory read of the inputs and write of the outputs.
//
// Reference semantics: karman-2d/karman_train.py:77-90 (to_feature/to_staggered), :173-185
// (KarmanFlow.step), :421-436 (correction add + loss); SURVEY.md Appendix A.
#include "sol_cells.cuh"
#include "sol_internal.cuh"

SOL_TRACE_TU()

namespace sol {

thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};
int g_pdl = 1;
// Measured on B200 at the bench shape (SOL-32, 3 simulations, 30 iterations, two runs each): fuse_small/fuse_solver_io =
// 1/0: 21.42 ms, 0/0: 20.9 ms, 1/1: 21.03 ms, 0/1: 20.39 ms per iteration.
int g_fuse_small = 0;        // fold corr_bwd (correction-gradient scaling) into diffuse_bc_bwd: one launch fewer, but slower
int g_fuse_solver_io = 1;
int g_deterministic = 0;      // ordered reductions instead of floating-point atomics where the engine has the choice (see sol_b200.h)    // fold to_feature / feat_bwd into the projection kernel (2 launches fewer per step)

static inline int grid_for(size_t n, int threads, int sm_count) {
    size_t blocks = (n + threads - 1) / threads;
    size_t cap = (size_t)sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// decode a flat face index of one simulation into (component, j, i)
struct FaceIdx { int comp, j, i; };
__device__ __forceinline__ FaceIdx decode_face(int rem, int NY, int X) {
    FaceIdx f;
    if (rem < NY) { f.comp = 0; f.j = rem / X; f.i = rem - f.j * X; }
    else { rem -= NY; f.comp = 1; f.j = rem / (X + 1); f.i = rem - f.j * (X + 1); }
    return f;
}

// ------------------------------------------------------------------------------------------------
// diffuse + BC (forward) and its adjoint
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_diffuse_bc(int B, int Y, int X, const float* __restrict__ re, float dt_res2,
                                                    const float* __restrict__ vy, const float* __restrict__ vx,
                                                    const float* __restrict__ bcm, const float* __restrict__ bcv,
                                                    float* __restrict__ vy_out, float* __restrict__ vx_out) {
    pdl_sync();
    const int NY = (Y + 1) * X, NX = Y * (X + 1), NF = NY + NX;
    const size_t total = (size_t)B * NF;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(idx / NF);
        const FaceIdx f = decode_face((int)(idx - (size_t)b * NF), NY, X);
        const float alpha = dt_res2 / re[b];
        if (f.comp == 0)
            vy_out[(size_t)b * NY + f.j * X + f.i] = diffuse_bc_cell(vy + (size_t)b * NY, Y + 1, X, f.j, f.i, alpha, bcm, bcv);
        else
            vx_out[(size_t)b * NX + f.j * (X + 1) + f.i] = diffuse_bc_cell(vx + (size_t)b * NX, Y, X + 1, f.j, f.i, alpha, nullptr, nullptr);
    }
}

__global__ void __launch_bounds__(256) k_diffuse_bc_bwd(int B, int Y, int X, const float* __restrict__ re, float dt_res2,
                                                        const float* __restrict__ gy, const float* __restrict__ gx,
                                                        const float* __restrict__ bcm,
                                                        const float* __restrict__ add_y, const float* __restrict__ add_x,
                                                        float* __restrict__ gy_in, float* __restrict__ gx_in,
                                                        float* __restrict__ g_corr, float sy, float sx,
                                                        float* __restrict__ zero_y, float* __restrict__ zero_x) {
    pdl_sync();
    const int NY = (Y + 1) * X, NX = Y * (X + 1), NF = NY + NX;
    const size_t total = (size_t)B * NF;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(idx / NF);
        const FaceIdx f = decode_face((int)(idx - (size_t)b * NF), NY, X);
        const float alpha = dt_res2 / re[b];
        if (f.comp == 0) {
            const size_t o = (size_t)b * NY + f.j * X + f.i;
            float v = diffuse_bc_bwd_cell(gy + (size_t)b * NY, Y + 1, X, f.j, f.i, alpha, bcm);
            if (add_y) v += add_y[o];
            gy_in[o] = v;
            if (zero_y) zero_y[o] = 0.0f;          // scatter target of the NEXT advection adjoint (no memset nodes in the sweep)
            // fused corr_bwd of the step that consumes this gradient: g_corr[b,j,i,0] = sigma_y * G_y[b,j,i]
            if (g_corr && f.j < Y) g_corr[((size_t)b * Y * X + f.j * X + f.i) * 2 + 0] = sy * v;
        } else {
            const size_t o = (size_t)b * NX + f.j * (X + 1) + f.i;
            float v = diffuse_bc_bwd_cell(gx + (size_t)b * NX, Y, X + 1, f.j, f.i, alpha, nullptr);
            if (add_x) v += add_x[o];
            gx_in[o] = v;
            if (zero_x) zero_x[o] = 0.0f;
            if (g_corr && f.i < X) g_corr[((size_t)b * Y * X + f.j * X + f.i) * 2 + 1] = sx * v;
        }
    }
}

int launch_diffuse_bc(const sol_plan* p, cudaStream_t st, int B, const float* re, float dt, float res,
                      const float* vy, const float* vx, float* vy_out, float* vx_out) {
    const size_t total = (size_t)B * (p->NY() + p->NX());
    SOL_CUDA(launch_kernel(k_diffuse_bc, dim3(grid_for(total, 256, p->sm_count)), dim3(256), 0, st, B, p->Y, p->X, re, dt * res * res, vy, vx, p->bc_mask_y,
                                                                    p->bc_val_y, vy_out, vx_out));
    SOL_LAUNCHED();
    return SOL_OK;
}

int launch_diffuse_bc_bwd(const sol_plan* p, cudaStream_t st, int B, const float* re, float dt, float res,
                          const float* gy, const float* gx, float* gy_in, float* gx_in, const float* add_y, const float* add_x,
                          float* g_corr, float sy, float sx, float* zero_y, float* zero_x) {
    const size_t total = (size_t)B * (p->NY() + p->NX());
    SOL_CUDA(launch_kernel(k_diffuse_bc_bwd, dim3(grid_for(total, 256, p->sm_count)), dim3(256), 0, st, B, p->Y, p->X, re, dt * res * res, gy, gx, p->bc_mask_y,
                                                                        add_y, add_x, gy_in, gx_in, g_corr, sy, sx, zero_y, zero_x));
    SOL_LAUNCHED();
    return SOL_OK;
}

// ------------------------------------------------------------------------------------------------
// semi-Lagrangian advection (velocity + optional density with inflow) and the velocity adjoint
// ------------------------------------------------------------------------------------------------
template <int WRAP>
__global__ void __launch_bounds__(256) k_advect(int B, int Y, int X, float s, float dt, const float* __restrict__ vy,
                                                const float* __restrict__ vx, const float* __restrict__ rho,
                                                const float* __restrict__ inflow, float* __restrict__ vy_out,
                                                float* __restrict__ vx_out, float* __restrict__ rho_out) {
    pdl_sync();
    const int NY = (Y + 1) * X, NX = Y * (X + 1), NC = Y * X;
    const int NF = NY + NX + (rho ? NC : 0);
    const size_t total = (size_t)B * NF;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(idx / NF);
        int rem = (int)(idx - (size_t)b * NF);
        const float* vyb = vy + (size_t)b * NY;
        const float* vxb = vx + (size_t)b * NX;
        if (rem < NY + NX) {
            const FaceIdx f = decode_face(rem, NY, X);
            if (f.comp == 0) vy_out[(size_t)b * NY + f.j * X + f.i] = advect_vy_cell<WRAP>(vyb, vxb, Y, X, f.j, f.i, s);
            else vx_out[(size_t)b * NX + f.j * (X + 1) + f.i] = advect_vx_cell<WRAP>(vyb, vxb, Y, X, f.j, f.i, s);
        } else {
            rem -= NY + NX;
            const int j = rem / X, i = rem - j * X;
            float r = advect_rho_cell(rho + (size_t)b * NC, vyb, vxb, Y, X, j, i, s);
            if (inflow) r += inflow[rem] * dt;
            rho_out[(size_t)b * NC + rem] = r;
        }
    }
}

template <int WRAP>
__global__ void __launch_bounds__(256) k_advect_bwd(int B, int Y, int X, float s, const float* __restrict__ vy,
                                                    const float* __restrict__ vx, const float* __restrict__ gy_out,
                                                    const float* __restrict__ gx_out, float* gy, float* gx) {
    pdl_sync();
    const int NY = (Y + 1) * X, NX = Y * (X + 1), NF = NY + NX;
    const size_t total = (size_t)B * NF;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(idx / NF);
        const FaceIdx f = decode_face((int)(idx - (size_t)b * NF), NY, X);
        const float* vyb = vy + (size_t)b * NY;
        const float* vxb = vx + (size_t)b * NX;
        float* gyb = gy + (size_t)b * NY;
        float* gxb = gx + (size_t)b * NX;
        if (f.comp == 0) {
            const float g = gy_out[(size_t)b * NY + f.j * X + f.i];
            if (g != 0.0f) advect_vy_cell_bwd<WRAP>(vyb, vxb, Y, X, f.j, f.i, s, g, gyb, gxb);
        } else {
            const float g = gx_out[(size_t)b * NX + f.j * (X + 1) + f.i];
            if (g != 0.0f) advect_vx_cell_bwd<WRAP>(vyb, vxb, Y, X, f.j, f.i, s, g, gyb, gxb);
        }
    }
}

int launch_advect(const sol_plan* p, cudaStream_t st, int B, float dt, const float* vy, const float* vx, const float* rho,
                  float* vy_out, float* vx_out, float* rho_out) {
    const size_t total = (size_t)B * (p->NY() + p->NX() + (rho ? p->NC() : 0));
    const float s = dt / p->dx;
    const int g = grid_for(total, 256, p->sm_count);
    if (p->boundary == SOL_BOUNDARY_PERIODIC)
        SOL_CUDA(launch_kernel(k_advect<WRAP_PERIODIC>, g, dim3(256), 0, st, B, p->Y, p->X, s, dt, vy, vx, nullptr, nullptr, vy_out, vx_out, nullptr));
    else
        SOL_CUDA(launch_kernel(k_advect<WRAP_REPLICATE>, g, dim3(256), 0, st, B, p->Y, p->X, s, dt, vy, vx, rho, p->inflow, vy_out, vx_out, rho_out));
    SOL_LAUNCHED();
    return SOL_OK;
}

int launch_advect_bwd(const sol_plan* p, cudaStream_t st, int B, float dt, const float* vy, const float* vx,
                      const float* gy_out, const float* gx_out, float* gy, float* gx, bool targets_are_zero) {
    if (!targets_are_zero) {
        SOL_CUDA(cudaMemsetAsync(gy, 0, (size_t)B * p->NY() * sizeof(float), st));
        SOL_CUDA(cudaMemsetAsync(gx, 0, (size_t)B * p->NX() * sizeof(float), st));
    }
    const size_t total = (size_t)B * (p->NY() + p->NX());
    const float s = dt / p->dx;
    const int g = grid_for(total, 256, p->sm_count);
    if (p->boundary == SOL_BOUNDARY_PERIODIC)
        SOL_CUDA(launch_kernel(k_advect_bwd<WRAP_PERIODIC>, g, dim3(256), 0, st, B, p->Y, p->X, s, vy, vx, gy_out, gx_out, gy, gx));
    else
        SOL_CUDA(launch_kernel(k_advect_bwd<WRAP_REPLICATE>, g, dim3(256), 0, st, B, p->Y, p->X, s, vy, vx, gy_out, gx_out, gy, gx));
    SOL_LAUNCHED();
    return SOL_OK;
}

// ------------------------------------------------------------------------------------------------
// stand-alone masked divergence
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_divergence(int B, int Y, int X, const float* __restrict__ vy, const float* __restrict__ vx,
                                                    const float* __restrict__ my, const float* __restrict__ mx, float* __restrict__ d) {
    pdl_sync();
    const int NY = (Y + 1) * X, NX = Y * (X + 1), NC = Y * X;
    const size_t total = (size_t)B * NC;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(idx / NC);
        const int rem = (int)(idx - (size_t)b * NC);
        const int j = rem / X, i = rem - j * X;
        d[idx] = divergence_cell(vy + (size_t)b * NY, vx + (size_t)b * NX, my, mx, Y, X, j, i);
    }
}

int launch_divergence(const sol_plan* p, cudaStream_t st, int B, const float* vy, const float* vx, float* div) {
    const size_t total = (size_t)B * p->NC();
    SOL_CUDA(launch_kernel(k_divergence, dim3(grid_for(total, 256, p->sm_count)), dim3(256), 0, st, B, p->Y, p->X, vy, vx, p->face_my, p->face_mx, div));
    SOL_LAUNCHED();
    return SOL_OK;
}

// ------------------------------------------------------------------------------------------------
// feature extraction, correction add + loss, and their adjoints
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_to_feature(int B, int Y, int X, const float* __restrict__ vy, const float* __restrict__ vx,
                                                    const float* __restrict__ re, float isy, float isx, float isr,
                                                    float* __restrict__ feat) {
    pdl_sync();
    const int NY = (Y + 1) * X, NX = Y * (X + 1), NC = Y * X;
    const size_t total = (size_t)B * NC;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(idx / NC);
        const int rem = (int)(idx - (size_t)b * NC);
        const int j = rem / X, i = rem - j * X;
        float* f = feat + idx * 3;
        f[0] = vy[(size_t)b * NY + j * X + i] * isy;
        f[1] = vx[(size_t)b * NX + j * (X + 1) + i] * isx;
        f[2] = re[b] * isr;
    }
}

int launch_to_feature(const sol_plan* p, cudaStream_t st, int B, const float* vy, const float* vx, const float* re,
                      float sy, float sx, float sr, float* feat) {
    const size_t total = (size_t)B * p->NC();
    SOL_CUDA(launch_kernel(k_to_feature, dim3(grid_for(total, 256, p->sm_count)), dim3(256), 0, st, B, p->Y, p->X, vy, vx, re, 1.0f / sy, 1.0f / sx, 1.0f / sr, feat));
    SOL_LAUNCHED();
    return SOL_OK;
}

// Burgers features (burgers_train.py:75-82, 398-415): [vy, vx (, fy, fx)][:Y,:X] / sigma, channel-last
__global__ void __launch_bounds__(256) k_to_feature_burgers(int B, int Y, int X, const float* __restrict__ vy, const float* __restrict__ vx,
                                                            const float* __restrict__ fy, const float* __restrict__ fx, float isy, float isx,
                                                            float isfy, float isfx, int cfeat, float* __restrict__ feat) {
    pdl_sync();
    const int NY = (Y + 1) * X, NX = Y * (X + 1), NC = Y * X;
    const size_t total = (size_t)B * NC;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(idx / NC);
        const int rem = (int)(idx - (size_t)b * NC);
        const int j = rem / X, i = rem - j * X;
        float* f = feat + idx * cfeat;
        const size_t oy = (size_t)b * NY + j * X + i, ox = (size_t)b * NX + j * (X + 1) + i;
        f[0] = vy[oy] * isy;
        f[1] = vx[ox] * isx;
        if (cfeat >= 4) {
            f[2] = fy[oy] * isfy;
            f[3] = fx[ox] * isfx;
        }
    }
}

int launch_to_feature_burgers(const sol_plan* p, cudaStream_t st, int B, const float* vy, const float* vx, const float* fy, const float* fx,
                              float sy, float sx, float sfy, float sfx, int cfeat, float* feat) {
    const size_t total = (size_t)B * p->NC();
    SOL_CUDA(launch_kernel(k_to_feature_burgers, dim3(grid_for(total, 256, p->sm_count)), dim3(256), 0, st, B, p->Y, p->X, vy, vx, fy, fx,
                           1.0f / sy, 1.0f / sx, 1.0f / sfy, 1.0f / sfx, cfeat, feat));
    SOL_LAUNCHED();
    return SOL_OK;
}

// G_out = G + add (the loss gradient of the previous step joins the state gradient); optionally also the fused corr_bwd
// of the step that consumes it: g_corr[b,j,i,:] = sigma * G_out[:Y,:X]
__global__ void __launch_bounds__(256) k_add_faces(int B, int Y, int X, const float* __restrict__ gy, const float* __restrict__ gx,
                                                   const float* __restrict__ add_y, const float* __restrict__ add_x,
                                                   float* __restrict__ gy_out, float* __restrict__ gx_out, float* __restrict__ g_corr,
                                                   float sy, float sx) {
    pdl_sync();
    const int NY = (Y + 1) * X, NX = Y * (X + 1), NF = NY + NX;
    const size_t total = (size_t)B * NF;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(idx / NF);
        const FaceIdx f = decode_face((int)(idx - (size_t)b * NF), NY, X);
        if (f.comp == 0) {
            const size_t o = (size_t)b * NY + f.j * X + f.i;
            const float v = gy[o] + (add_y ? add_y[o] : 0.0f);
            gy_out[o] = v;
            if (g_corr && f.j < Y) g_corr[((size_t)b * Y * X + f.j * X + f.i) * 2 + 0] = sy * v;
        } else {
            const size_t o = (size_t)b * NX + f.j * (X + 1) + f.i;
            const float v = gx[o] + (add_x ? add_x[o] : 0.0f);
            gx_out[o] = v;
            if (g_corr && f.i < X) g_corr[((size_t)b * Y * X + f.j * X + f.i) * 2 + 1] = sx * v;
        }
    }
}

int launch_add_faces(const sol_plan* p, cudaStream_t st, int B, const float* gy, const float* gx, const float* add_y, const float* add_x,
                     float* gy_out, float* gx_out, float* g_corr, float sy, float sx) {
    const size_t total = (size_t)B * (p->NY() + p->NX());
    SOL_CUDA(launch_kernel(k_add_faces, dim3(grid_for(total, 256, p->sm_count)), dim3(256), 0, st, B, p->Y, p->X, gy, gx, add_y, add_x, gy_out,
                           gx_out, g_corr, sy, sx));
    SOL_LAUNCHED();
    return SOL_OK;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// v_out = v + sigma*corr on the [:Y,:X] block; loss_i += 1/2 ((gt - v_out)/sigma)^2;
// gl = (v_out - gt)/sigma^2 * inv_m     (karman_train.py:421-436)
__global__ void __launch_bounds__(256) k_correct_loss(int B, int Y, int X, const float* __restrict__ vy, const float* __restrict__ vx,
                                                      const float* __restrict__ corr, float sy, float sx,
                                                      const float* __restrict__ gt_vy, const float* __restrict__ gt_vx, float inv_m,
                                                      float* __restrict__ vy_out, float* __restrict__ vx_out,
                                                      float* __restrict__ gl_vy, float* __restrict__ gl_vx, float* loss,
                                                      float* __restrict__ loss_part) {
    pdl_sync();
    const int NY = (Y + 1) * X, NX = Y * (X + 1), NF = NY + NX, NC = Y * X;
    const size_t total = (size_t)B * NF;
    float part = 0.0f;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(idx / NF);
        const FaceIdx f = decode_face((int)(idx - (size_t)b * NF), NY, X);
        if (f.comp == 0) {
            const size_t o = (size_t)b * NY + f.j * X + f.i;
            float v = vy[o];
            if (corr && f.j < Y) v += sy * corr[((size_t)b * NC + f.j * X + f.i) * 2 + 0];
            vy_out[o] = v;
            if (gt_vy) {
                const float dlt = (v - gt_vy[o]) / sy;
                part += 0.5f * dlt * dlt;
                if (gl_vy) gl_vy[o] = dlt / sy * inv_m;
            }
        } else {
            const size_t o = (size_t)b * NX + f.j * (X + 1) + f.i;
            float v = vx[o];
            if (corr && f.i < X) v += sx * corr[((size_t)b * NC + f.j * X + f.i) * 2 + 1];
            vx_out[o] = v;
            if (gt_vx) {
                const float dlt = (v - gt_vx[o]) / sx;
                part += 0.5f * dlt * dlt;
                if (gl_vx) gl_vx[o] = dlt / sx * inv_m;
            }
        }
    }
    if (loss) {
        __shared__ float red[8];
        part = warp_sum(part);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
        __syncthreads();
        if (threadIdx.x < 32) {
            float v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0f;
            v = warp_sum(v);
            // deterministic mode: one partial per CTA, summed in CTA order by k_loss_finalize
            if (threadIdx.x == 0) { if (loss_part) loss_part[blockIdx.x] = v; else atomicAdd(loss, v); }
        }
    }
}

int launch_correct_loss(const sol_plan* p, cudaStream_t st, int B, const float* vy, const float* vx, const float* corr,
                        float sy, float sx, const float* gt_vy, const float* gt_vx, float inv_m,
                        float* vy_out, float* vx_out, float* gl_vy, float* gl_vx, float* loss, float* loss_part) {
    const int g = correct_loss_grid(p, B);
    SOL_CUDA(launch_kernel(k_correct_loss, g, dim3(256), 0, st, B, p->Y, p->X, vy, vx, corr, sy, sx, gt_vy, gt_vx, inv_m, vy_out, vx_out, gl_vy, gl_vx,
                                      gt_vy ? loss : nullptr, gt_vy ? loss_part : nullptr));
    SOL_LAUNCHED();
    return SOL_OK;
}

int correct_loss_grid(const sol_plan* p, int B) {
    const size_t total = (size_t)B * (p->NY() + p->NX());
    int g = grid_for(total, 256, p->sm_count);
    if (g > p->sm_count) g = p->sm_count;   // few atomics on the loss scalar / few partials
    return g;
}

__global__ void __launch_bounds__(32) k_loss_finalize(const float* __restrict__ part, int stride, int nparts, float* __restrict__ loss) {
    pdl_sync();
    if (threadIdx.x != 0) return;
    float s = 0.0f;
    for (int c = 0; c < nparts; ++c) s += part[(size_t)blockIdx.x * stride + c];      // fixed order
    loss[blockIdx.x] = s;
}

int launch_loss_finalize(cudaStream_t st, int msteps, const float* loss_part, int stride, float* loss) {
    SOL_CUDA(launch_kernel(k_loss_finalize, dim3(msteps), dim3(32), 0, st, loss_part, stride, stride, loss));
    SOL_LAUNCHED();
    return SOL_OK;
}

__global__ void __launch_bounds__(256) k_corr_bwd(int B, int Y, int X, const float* __restrict__ Gy, const float* __restrict__ Gx,
                                                  float sy, float sx, float* __restrict__ g_corr) {
    pdl_sync();
    const int NY = (Y + 1) * X, NX = Y * (X + 1), NC = Y * X;
    const size_t total = (size_t)B * NC;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(idx / NC);
        const int rem = (int)(idx - (size_t)b * NC);
        const int j = rem / X, i = rem - j * X;
        float2 o;
        o.x = sy * Gy[(size_t)b * NY + j * X + i];
        o.y = sx * Gx[(size_t)b * NX + j * (X + 1) + i];
        reinterpret_cast<float2*>(g_corr)[idx] = o;
    }
}

int launch_corr_bwd(const sol_plan* p, cudaStream_t st, int B, const float* Gy, const float* Gx, float sy, float sx, float* g_corr) {
    const size_t total = (size_t)B * p->NC();
    SOL_CUDA(launch_kernel(k_corr_bwd, dim3(grid_for(total, 256, p->sm_count)), dim3(256), 0, st, B, p->Y, p->X, Gy, Gx, sy, sx, g_corr));
    SOL_LAUNCHED();
    return SOL_OK;
}

__global__ void __launch_bounds__(256) k_feat_bwd(int B, int Y, int X, const float* __restrict__ Gy, const float* __restrict__ Gx,
                                                  const float* __restrict__ g_feat, int cfeat, float isy, float isx,
                                                  float* __restrict__ Gy_out, float* __restrict__ Gx_out) {
    pdl_sync();
    const int NY = (Y + 1) * X, NX = Y * (X + 1), NF = NY + NX, NC = Y * X;
    const size_t total = (size_t)B * NF;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(idx / NF);
        const FaceIdx f = decode_face((int)(idx - (size_t)b * NF), NY, X);
        if (f.comp == 0) {
            const size_t o = (size_t)b * NY + f.j * X + f.i;
            float v = Gy[o];
            if (f.j < Y) v += g_feat[((size_t)b * NC + f.j * X + f.i) * cfeat + 0] * isy;
            Gy_out[o] = v;
        } else {
            const size_t o = (size_t)b * NX + f.j * (X + 1) + f.i;
            float v = Gx[o];
            if (f.i < X) v += g_feat[((size_t)b * NC + f.j * X + f.i) * cfeat + 1] * isx;
            Gx_out[o] = v;
        }
    }
}

int launch_feat_bwd(const sol_plan* p, cudaStream_t st, int B, const float* Gy, const float* Gx, const float* g_feat, int cfeat,
                    float sy, float sx, float* Gy_out, float* Gx_out) {
    const size_t total = (size_t)B * (p->NY() + p->NX());
    SOL_CUDA(launch_kernel(k_feat_bwd, dim3(grid_for(total, 256, p->sm_count)), dim3(256), 0, st, B, p->Y, p->X, Gy, Gx, g_feat, cfeat, 1.0f / sy, 1.0f / sx, Gy_out, Gx_out));
    SOL_LAUNCHED();
    return SOL_OK;
}

// ------------------------------------------------------------------------------------------------
// Burgers diffusion: periodic operator applied as a dense circular convolution with the
// real-space kernel of exp(-(2 pi k)^2 nu dt) (or explicit 5-point when no kernel is given),
// fused with the forcing term  + dt*f   (burgers_train.py:183-187)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_burgers_diffuse(int H, int W, size_t stride, float amount, const float* __restrict__ ker,
                                                         const float* __restrict__ v, const float* __restrict__ f, float dtf,
                                                         float* __restrict__ out) {
    pdl_sync();
    extern __shared__ float sm[];
    float* sv = sm;            // [H*W]
    float* sk = sm + H * W;    // [H*W]
    const int b = blockIdx.x;
    const int n = H * W;
    const float* vb = v + (size_t)b * stride;
    for (int k = threadIdx.x; k < n; k += blockDim.x) { sv[k] = vb[k]; if (ker) sk[k] = ker[k]; }
    __syncthreads();
    for (int o = threadIdx.x; o < n; o += blockDim.x) {
        const int j = o / W, i = o - j * W;
        float acc;
        if (ker) {
            acc = 0.0f;
            for (int a = 0; a < H; ++a) {
                int jj = j - a; if (jj < 0) jj += H;
                const float* kr = sk + a * W;
                const float* vr = sv + jj * W;
                for (int c = 0; c < W; ++c) {
                    int ii = i - c; if (ii < 0) ii += W;
                    acc = fmaf(kr[c], vr[ii], acc);
                }
            }
        } else {
            acc = sv[o] + amount * lap5<WRAP_PERIODIC>(sv, H, W, j, i);
        }
        if (f) acc += dtf * f[(size_t)b * stride + o];
        out[(size_t)b * stride + o] = acc;
    }
}

int launch_burgers_diffuse(const sol_plan* p, cudaStream_t st, int B, float amount, const float* ky, const float* kx,
                           const float* vy, const float* vx, const float* fy, const float* fx, float dtf,
                           float* vy_out, float* vx_out) {
    const size_t smy = 2 * p->NY() * sizeof(float), smx = 2 * p->NX() * sizeof(float);
    if (smy > 200 * 1024 || smx > 200 * 1024) return fail(SOL_ERR_UNSUPPORTED, "burgers diffusion: grid too large for the shared-memory kernel");
    static bool attr_set = false;
    if (!attr_set) {
        SOL_CUDA(cudaFuncSetAttribute(k_burgers_diffuse, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    SOL_CUDA(launch_kernel(k_burgers_diffuse, dim3(B), dim3(256), smy, st, p->Y + 1, p->X, p->NY(), amount, ky, vy, fy, dtf, vy_out));
    SOL_LAUNCHED();
    SOL_CUDA(launch_kernel(k_burgers_diffuse, dim3(B), dim3(256), smx, st, p->Y, p->X + 1, p->NX(), amount, kx, vx, fx, dtf, vx_out));
    SOL_LAUNCHED();
    return SOL_OK;
}

// ------------------------------------------------------------------------------------------------
// TF1 Adam (karman_train.py:449): theta -= lr_t * m / (sqrt(v) + eps)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_adam(size_t n, float* __restrict__ theta, const float* __restrict__ g, float* __restrict__ m,
                                              float* __restrict__ v, float lr_t, float b1, float b2, float eps, float gscale) {
    pdl_sync();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float gi = g[i] * gscale;
        const float mi = b1 * m[i] + (1.0f - b1) * gi;
        const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        theta[i] -= lr_t * mi / (sqrtf(vi) + eps);
    }
}

int launch_adam(cudaStream_t st, size_t n, float* theta, const float* g, float* m, float* v, float lr_t, float b1, float b2,
                float eps, float gscale) {
    SOL_CUDA(launch_kernel(k_adam, dim3(grid_for(n, 256, 148)), dim3(256), 0, st, n, theta, g, m, v, lr_t, b1, b2, eps, gscale));
    SOL_LAUNCHED();
    return SOL_OK;
}

}  // namespace sol
