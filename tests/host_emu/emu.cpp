// TEST INFRASTRUCTURE ONLY.  Compiles the per-cell stage functions of
// solver_in_the_loop_b200/csrc/sol_cells.cuh with g++ so that the -m "not gpu" test-suite can
// compare the exact arithmetic the CUDA kernels run per cell with the CPU oracle in a container
// that has no GPU.  Nothing in the product imports or links this file.
#include "../../solver_in_the_loop_b200/csrc/sol_cells.cuh"

using namespace sol;

extern "C" {

void emu_diffuse_bc(int B, int Y, int X, const float* re, float dt_res2, const float* vy, const float* vx, const float* bcm,
                    const float* bcv, float* vy_out, float* vx_out) {
    const int NY = (Y + 1) * X, NX = Y * (X + 1);
    for (int b = 0; b < B; ++b) {
        const float alpha = dt_res2 / re[b];
        for (int j = 0; j <= Y; ++j)
            for (int i = 0; i < X; ++i) vy_out[b * NY + j * X + i] = diffuse_bc_cell(vy + b * NY, Y + 1, X, j, i, alpha, bcm, bcv);
        for (int j = 0; j < Y; ++j)
            for (int i = 0; i <= X; ++i) vx_out[b * NX + j * (X + 1) + i] = diffuse_bc_cell(vx + b * NX, Y, X + 1, j, i, alpha, nullptr, nullptr);
    }
}

void emu_diffuse_bc_bwd(int B, int Y, int X, const float* re, float dt_res2, const float* gy, const float* gx, const float* bcm,
                        float* gy_in, float* gx_in) {
    const int NY = (Y + 1) * X, NX = Y * (X + 1);
    for (int b = 0; b < B; ++b) {
        const float alpha = dt_res2 / re[b];
        for (int j = 0; j <= Y; ++j)
            for (int i = 0; i < X; ++i) gy_in[b * NY + j * X + i] = diffuse_bc_bwd_cell(gy + b * NY, Y + 1, X, j, i, alpha, bcm);
        for (int j = 0; j < Y; ++j)
            for (int i = 0; i <= X; ++i) gx_in[b * NX + j * (X + 1) + i] = diffuse_bc_bwd_cell(gx + b * NX, Y, X + 1, j, i, alpha, nullptr);
    }
}

void emu_advect(int B, int Y, int X, float s, float dt, int periodic, const float* vy, const float* vx, const float* rho,
                const float* inflow, float* vy_out, float* vx_out, float* rho_out) {
    const int NY = (Y + 1) * X, NX = Y * (X + 1), NC = Y * X;
    for (int b = 0; b < B; ++b) {
        const float* vyb = vy + b * NY; const float* vxb = vx + b * NX;
        for (int j = 0; j <= Y; ++j)
            for (int i = 0; i < X; ++i)
                vy_out[b * NY + j * X + i] = periodic ? advect_vy_cell<WRAP_PERIODIC>(vyb, vxb, Y, X, j, i, s)
                                                      : advect_vy_cell<WRAP_REPLICATE>(vyb, vxb, Y, X, j, i, s);
        for (int j = 0; j < Y; ++j)
            for (int i = 0; i <= X; ++i)
                vx_out[b * NX + j * (X + 1) + i] = periodic ? advect_vx_cell<WRAP_PERIODIC>(vyb, vxb, Y, X, j, i, s)
                                                            : advect_vx_cell<WRAP_REPLICATE>(vyb, vxb, Y, X, j, i, s);
        if (rho)
            for (int j = 0; j < Y; ++j)
                for (int i = 0; i < X; ++i) {
                    float r = advect_rho_cell(rho + b * NC, vyb, vxb, Y, X, j, i, s);
                    if (inflow) r += inflow[j * X + i] * dt;
                    rho_out[b * NC + j * X + i] = r;
                }
    }
}

void emu_advect_bwd(int B, int Y, int X, float s, int periodic, const float* vy, const float* vx, const float* gy_out,
                    const float* gx_out, float* gy, float* gx) {
    const int NY = (Y + 1) * X, NX = Y * (X + 1);
    for (int k = 0; k < B * NY; ++k) gy[k] = 0.f;
    for (int k = 0; k < B * NX; ++k) gx[k] = 0.f;
    for (int b = 0; b < B; ++b) {
        const float* vyb = vy + b * NY; const float* vxb = vx + b * NX;
        float* gyb = gy + b * NY; float* gxb = gx + b * NX;
        for (int j = 0; j <= Y; ++j)
            for (int i = 0; i < X; ++i) {
                const float g = gy_out[b * NY + j * X + i];
                if (periodic) advect_vy_cell_bwd<WRAP_PERIODIC>(vyb, vxb, Y, X, j, i, s, g, gyb, gxb);
                else advect_vy_cell_bwd<WRAP_REPLICATE>(vyb, vxb, Y, X, j, i, s, g, gyb, gxb);
            }
        for (int j = 0; j < Y; ++j)
            for (int i = 0; i <= X; ++i) {
                const float g = gx_out[b * NX + j * (X + 1) + i];
                if (periodic) advect_vx_cell_bwd<WRAP_PERIODIC>(vyb, vxb, Y, X, j, i, s, g, gyb, gxb);
                else advect_vx_cell_bwd<WRAP_REPLICATE>(vyb, vxb, Y, X, j, i, s, g, gyb, gxb);
            }
    }
}

void emu_divergence(int B, int Y, int X, const float* vy, const float* vx, const float* my, const float* mx, float* d) {
    const int NY = (Y + 1) * X, NX = Y * (X + 1), NC = Y * X;
    for (int b = 0; b < B; ++b)
        for (int j = 0; j < Y; ++j)
            for (int i = 0; i < X; ++i) d[b * NC + j * X + i] = divergence_cell(vy + b * NY, vx + b * NX, my, mx, Y, X, j, i);
}

void emu_gradsub(int B, int Y, int X, const float* vy, const float* vx, const float* p, const float* my, const float* mx,
                 float* vy_out, float* vx_out) {
    const int NY = (Y + 1) * X, NX = Y * (X + 1), NC = Y * X;
    for (int b = 0; b < B; ++b) {
        for (int j = 0; j <= Y; ++j)
            for (int i = 0; i < X; ++i) vy_out[b * NY + j * X + i] = gradsub_vy_cell(vy + b * NY, p + b * NC, my, Y, X, j, i);
        for (int j = 0; j < Y; ++j)
            for (int i = 0; i <= X; ++i) vx_out[b * NX + j * (X + 1) + i] = gradsub_vx_cell(vx + b * NX, p + b * NC, mx, Y, X, j, i);
    }
}

void emu_laplace(int B, int Y, int X, const float* p, const unsigned char* active, const float* diag, float* out) {
    const int NC = Y * X;
    for (int b = 0; b < B; ++b)
        for (int j = 0; j < Y; ++j)
            for (int i = 0; i < X; ++i) out[b * NC + j * X + i] = laplace_cell(p + b * NC, active, diag, Y, X, j, i);
}

}  // extern "C"

// ---- direct pressure projection (sol_direct_host.h + the arithmetic of k_direct_solve / k_direct_apply in fp32) ----
#include "../../solver_in_the_loop_b200/csrc/sol_direct_host.h"

extern "C" int emu_direct_solve(int Y, int X, int B, const unsigned char* act, const float* diag, const float* rhs, float* p_out) {
    DirectHost h;
    if (!direct_precompute(Y, X, act, diag, h)) return -1;
    const int N = Y * X;
    std::vector<float> D(N), U(N), V(N), Z(N), P0(N), s(h.kp);
    for (int b = 0; b < B; ++b) {
        const float* r = rhs + (size_t)b * N;
        for (int c = 0; c < N; ++c) D[c] = act[c] ? r[c] : 0.0f;
        // U = Sy D
        for (int a = 0; a < Y; ++a)
            for (int i = 0; i < X; ++i) { float acc = 0.f; for (int m = 0; m < Y; ++m) acc = fmaf(h.Sy[(size_t)m * Y + a], D[(size_t)m * X + i], acc); U[(size_t)a * X + i] = acc; }
        // V = (U Sx) * ilam
        for (int a = 0; a < Y; ++a)
            for (int i = 0; i < X; ++i) { float acc = 0.f; for (int m = 0; m < X; ++m) acc = fmaf(U[(size_t)a * X + m], h.Sx[(size_t)m * X + i], acc); V[(size_t)a * X + i] = acc * h.ilam[(size_t)a * X + i]; }
        // Z = V Sx
        for (int a = 0; a < Y; ++a)
            for (int i = 0; i < X; ++i) { float acc = 0.f; for (int m = 0; m < X; ++m) acc = fmaf(V[(size_t)a * X + m], h.Sx[(size_t)m * X + i], acc); Z[(size_t)a * X + i] = acc; }
        // p0 = Sy Z
        for (int a = 0; a < Y; ++a)
            for (int i = 0; i < X; ++i) { float acc = 0.f; for (int m = 0; m < Y; ++m) acc = fmaf(h.Sy[(size_t)m * Y + a], Z[(size_t)m * X + i], acc); P0[(size_t)a * X + i] = acc; }
        for (int q = 0; q < h.kp; ++q) {
            float sv = 0.f;
            if (q < h.k)
                for (int e = 0; e < 5; ++e) { const int col = h.rt_col[q * 5 + e]; if (col >= 0) sv = fmaf(h.rt_val[q * 5 + e], P0[col], sv); }
            s[q] = sv;
        }
        for (int c = 0; c < N; ++c) {
            float corr = 0.f;
            for (int q = 0; q < h.k; ++q) corr = fmaf(h.Wt[(size_t)q * N + c], s[q], corr);
            p_out[(size_t)b * N + c] = act[c] ? P0[c] - corr : -r[c] / diag[c];
        }
    }
    return h.k;
}
