// Host-side precomputation of the direct pressure projection (see sol_direct.cu): plain C++, no CUDA — shared by the library
// (sol_direct.cu uploads the result) and by the host-emulation tests (tests/host_emu/emu.cpp, test infrastructure).
//
//   A0 = 5-point Laplacian of the Y x X rectangle with p = 0 one cell outside  =  S (Lambda) S,  S = type-I sine transform
//   A' = A0 + E R^T   (k rows changed by the obstacle)
//   A'^-1 d = p0 - (W M) (R^T p0),   p0 = A0^-1 d,   W = A0^-1 E,   M = (I + R^T W)^-1
// A' is the scene's operator A on the fluid cells.  On the solid cells A is diagonal (p = -d/diag, which the apply kernel writes
// itself); A' keeps the A0 stencil INSIDE the solid region and only cuts the fluid <-> solid links, so the solid block stays
// decoupled from the fluid one (the fluid pressures are those of A) and only the cells on either side of the obstacle surface
// change: k ~ 2 x perimeter (76 at 128x64, ~150 at 256x128) instead of area + perimeter (164 / ~600).  The correction basis
// W M [N x k] — the dominant memory stream of the projection — shrinks by the same factor.
#pragma once
#include <math.h>

#include <algorithm>
#include <vector>

namespace sol {

struct DirectHost {
    int k = 0, kp = 0;                  // changed rows, padded to a multiple of 32
    std::vector<float> Sy, Sx, ilam;    // [Y*Y], [X*X], [Y*X] (1/lambda)
    std::vector<int> rt_col;            // [kp*5] sparse rows of R^T (-1 = unused entry)
    std::vector<float> rt_val;          // [kp*5]
    std::vector<float> Wt;              // [kp*N]: Wt[q][c] = (W M)[c][q], zero rows for q >= k
};

// act[c] != 0: fluid cell; dg[c]: number of accessible neighbours (>= 1), the negated diagonal of A.  Returns false when the
// capacitance matrix is singular.
inline bool direct_precompute(int Y, int X, const unsigned char* act, const float* dg, DirectHost& out) {
    const int N = Y * X;
    const double PI = 3.14159265358979323846;
    std::vector<double> Sy((size_t)Y * Y), Sx((size_t)X * X), il((size_t)N);
    for (int a = 0; a < Y; ++a)
        for (int b = 0; b < Y; ++b) Sy[(size_t)a * Y + b] = sqrt(2.0 / (Y + 1)) * sin(PI * (a + 1) * (b + 1) / (Y + 1));
    for (int a = 0; a < X; ++a)
        for (int b = 0; b < X; ++b) Sx[(size_t)a * X + b] = sqrt(2.0 / (X + 1)) * sin(PI * (a + 1) * (b + 1) / (X + 1));
    for (int a = 0; a < Y; ++a)
        for (int b = 0; b < X; ++b) il[(size_t)a * X + b] = 1.0 / (-4.0 + 2.0 * cos(PI * (a + 1) / (Y + 1)) + 2.0 * cos(PI * (b + 1) / (X + 1)));
    // rows of A' that differ from A0: cells with an in-domain neighbour of the other kind (fluid next to solid, solid next to fluid)
    auto inside = [&](int j, int i) -> bool { return j >= 0 && j < Y && i >= 0 && i < X; };
    auto other = [&](int c, int j, int i) -> bool { return inside(j, i) && (act[(size_t)j * X + i] != 0) != (act[c] != 0); };
    std::vector<int> rows;
    for (int j = 0; j < Y; ++j)
        for (int i = 0; i < X; ++i) {
            const int c = j * X + i;
            if (other(c, j - 1, i) || other(c, j + 1, i) || other(c, j, i - 1) || other(c, j, i + 1)) rows.push_back(c);
        }
    const int k = (int)rows.size();
    const int kp = k == 0 ? 32 : (k + 31) / 32 * 32;
    std::vector<int> rt_col((size_t)kp * 5, -1);
    std::vector<float> rt_val((size_t)kp * 5, 0.0f);
    std::vector<double> rtv((size_t)kp * 5, 0.0);
    for (int q = 0; q < k; ++q) {
        const int c = rows[q], j = c / X, i = c - j * X;
        // fluid row of A': -diag[c] on the diagonal (the accessible neighbours), +1 to fluid in-domain neighbours; solid row of A':
        // the A0 row without its links to fluid cells;  A0 row: -4, +1 to in-domain neighbours
        rt_col[q * 5] = c; rtv[q * 5] = act[c] ? 4.0 - (double)dg[c] : 0.0;
        const int nb[4][2] = {{j - 1, i}, {j + 1, i}, {j, i - 1}, {j, i + 1}};
        for (int e = 0; e < 4; ++e) {
            const int jj = nb[e][0], ii = nb[e][1];
            if (other(c, jj, ii)) { rt_col[q * 5 + 1 + e] = jj * X + ii; rtv[q * 5 + 1 + e] = -1.0; }
        }
    }
    for (size_t e = 0; e < rtv.size(); ++e) rt_val[e] = (float)rtv[e];
    // W[:, q] = A0^-1 e_c = Sy ((Sy[:,j] (x) Sx[:,i]) * ilam) Sx
    std::vector<double> W((size_t)N * (k > 0 ? k : 1), 0.0), G((size_t)N), T((size_t)N);
    for (int q = 0; q < k; ++q) {
        const int c = rows[q], j = c / X, i = c - j * X;
        for (int a = 0; a < Y; ++a)
            for (int b = 0; b < X; ++b) G[(size_t)a * X + b] = Sy[(size_t)a * Y + j] * Sx[(size_t)b * X + i] * il[(size_t)a * X + b];
        // T = G Sx
        for (int a = 0; a < Y; ++a) {
            double* tr = &T[(size_t)a * X];
            for (int b = 0; b < X; ++b) tr[b] = 0.0;
            for (int m = 0; m < X; ++m) {
                const double g = G[(size_t)a * X + m];
                const double* sr = &Sx[(size_t)m * X];
                for (int b = 0; b < X; ++b) tr[b] += g * sr[b];
            }
        }
        // W[:, q] = Sy T
        for (int a = 0; a < Y; ++a) {
            double out[256];
            for (int b = 0; b < X; ++b) out[b] = 0.0;
            for (int m = 0; m < Y; ++m) {
                const double s = Sy[(size_t)a * Y + m];
                const double* tr = &T[(size_t)m * X];
                for (int b = 0; b < X; ++b) out[b] += s * tr[b];
            }
            for (int b = 0; b < X; ++b) W[((size_t)a * X + b) * k + q] = out[b];
        }
    }
    // M = (I + R^T W)^-1 by Gauss-Jordan with partial pivoting
    std::vector<double> C((size_t)k * k, 0.0), Inv((size_t)k * k, 0.0);
    for (int q = 0; q < k; ++q) {
        Inv[(size_t)q * k + q] = 1.0;
        for (int e = 0; e < 5; ++e) {
            const int col = rt_col[q * 5 + e];
            if (col < 0) continue;
            const double v = rtv[q * 5 + e];
            for (int r = 0; r < k; ++r) C[(size_t)q * k + r] += v * W[(size_t)col * k + r];
        }
        C[(size_t)q * k + q] += 1.0;
    }
    for (int col = 0; col < k; ++col) {
        int piv = col;
        for (int r = col + 1; r < k; ++r)
            if (fabs(C[(size_t)r * k + col]) > fabs(C[(size_t)piv * k + col])) piv = r;
        if (fabs(C[(size_t)piv * k + col]) < 1e-13) return false;      // singular capacitance matrix: no direct solver for this scene
        if (piv != col)
            for (int m = 0; m < k; ++m) { std::swap(C[(size_t)piv * k + m], C[(size_t)col * k + m]); std::swap(Inv[(size_t)piv * k + m], Inv[(size_t)col * k + m]); }
        const double dv = C[(size_t)col * k + col];
        for (int m = 0; m < k; ++m) { C[(size_t)col * k + m] /= dv; Inv[(size_t)col * k + m] /= dv; }
        for (int r = 0; r < k; ++r) {
            if (r == col) continue;
            const double f = C[(size_t)r * k + col];
            if (f == 0.0) continue;
            for (int m = 0; m < k; ++m) { C[(size_t)r * k + m] -= f * C[(size_t)col * k + m]; Inv[(size_t)r * k + m] -= f * Inv[(size_t)col * k + m]; }
        }
    }
    // fold the capacitance matrix into the correction basis: (W M)[c][q] = sum_e W[c][e] M[e][q]
    std::vector<float> Wt((size_t)kp * N, 0.0f);
    {
        std::vector<double> row((size_t)(k > 0 ? k : 1));
        for (int c = 0; c < N; ++c) {
            for (int q = 0; q < k; ++q) row[q] = 0.0;
            for (int e = 0; e < k; ++e) {
                const double w = W[(size_t)c * k + e];
                if (w == 0.0) continue;
                const double* mr = &Inv[(size_t)e * k];
                for (int q = 0; q < k; ++q) row[q] += w * mr[q];
            }
            for (int q = 0; q < k; ++q) Wt[(size_t)q * N + c] = (float)row[q];
        }
    }
    out.k = k; out.kp = kp;
    out.Sy.assign(Sy.begin(), Sy.end()); out.Sx.assign(Sx.begin(), Sx.end()); out.ilam.assign(il.begin(), il.end());
    out.rt_col = rt_col; out.rt_val = rt_val; out.Wt = Wt;
    return true;
}

}  // namespace sol
