// Small constant tables derived numerically on the host from the element routines
// (no hand-typed stencils): Jacobi-diagonal coefficients, P2->P2 nested interpolation
// weights and the moment-coarsening maps of the Galerkin multigrid hierarchy.
// Plain C++ (no CUDA) so tests/hostcheck can validate them on the CPU.
#pragma once

#include <cstring>

#include "tm_element.cuh"

namespace tmx {

struct DiagTable {
    double Q[2][6][2][6];  // [triangle type][local node][component][moment slot]
};

// diag(K_T)(dof) = sum_m Q[..][m] w[m]
inline DiagTable make_diag_table(const Material<double>& mat) {
    DiagTable t;
    for (int type = 0; type < 2; ++type)
        for (int kl = 0; kl < 6; ++kl)
            for (int c = 0; c < 2; ++c)
                for (int m = 0; m < 6; ++m) {
                    double U[3][2] = {}, M[3][2] = {}, YU[3][2] = {}, YM[3][2] = {}, w[6] = {};
                    w[m] = 1.0;
                    if (kl < 3) U[kl][c] = 1.0; else M[kl - 3][c] = 1.0;
                    if (type == 0) tri_apply<double, false>(U, M, w, mat, YU, YM);
                    else tri_apply<double, true>(U, M, w, mat, YU, YM);
                    t.Q[type][kl][c][m] = kl < 3 ? YU[kl][c] : YM[kl - 3][c];
                }
    return t;
}

// Nested P2 interpolation coarse -> fine.  A coarse cell spans 4x4 fine lattice steps; a fine
// node at local position (a,b) in 0..4 takes  sum_q Pw[5b+a][q] * coarse cell-local node q.
struct TransferTable {
    double Pw[25][9];
    // restriction as a gather: weight of the coarse node of parity type (I&1) + 2*(J&1) at the
    // fine node offset (di, dj) in -3..3 from it (dense 7x7, mostly zeros)
    double Rw[4][49];
};

inline void coarse_barycentrics(int type, double x, double y, double lam[3]) {
    if (type == 0) { lam[0] = 1 - x; lam[1] = x - y; lam[2] = y; }
    else { lam[0] = 1 - y; lam[1] = y - x; lam[2] = x; }
}

inline TransferTable make_transfer_table() {
    TransferTable t;
    std::memset(&t, 0, sizeof(t));
    for (int b = 0; b <= 4; ++b)
        for (int a = 0; a <= 4; ++a) {
            const double x = a / 4.0, y = b / 4.0;
            const int type = (x >= y) ? 0 : 1;
            double l[3];
            coarse_barycentrics(type, x, y, l);
            const double phi[6] = {l[0] * (2 * l[0] - 1), l[1] * (2 * l[1] - 1), l[2] * (2 * l[2] - 1),
                                   4 * l[0] * l[1], 4 * l[1] * l[2], 4 * l[0] * l[2]};
            for (int k = 0; k < 6; ++k) t.Pw[5 * b + a][tri_local_to_cell(type, k)] += phi[k];
        }
    // translation invariant: evaluate around a coarse node well inside a large mesh
    for (int type = 0; type < 4; ++type) {
        const int I = 8 + (type & 1), J = 8 + (type >> 1);
        for (int dj = -3; dj <= 3; ++dj)
            for (int di = -3; di <= 3; ++di) {
                const int i = 2 * I + di, j = 2 * J + dj;
                const int cx = i >> 2, cy = j >> 2;
                const int qx = I - 2 * cx, qy = J - 2 * cy;
                double w = 0.0;
                if (qx >= 0 && qx <= 2 && qy >= 0 && qy <= 2) w = t.Pw[5 * (j - 4 * cy) + (i - 4 * cx)][3 * qy + qx];
                t.Rw[type][7 * (dj + 3) + (di + 3)] = w;
            }
    }
    return t;
}

// Coarse triangle = 4 fine triangles.  Entry s (0..3 for coarse T_A, 4..7 for coarse T_B):
// fine cell offset (dx,dy) in the 2x2 block, fine triangle type, and M[a][c] = coarse
// barycentric a at fine vertex c, so that  w^c = 1/4 sum_s M w^s M^T  (w normalised by 2|T|).
struct CoarsenTable {
    int dx[8], dy[8], type[8];
    double M[8][3][3];
};

inline CoarsenTable make_coarsen_table() {
    CoarsenTable t;
    const int spec[8][3] = {{0, 0, 0}, {0, 1, 0}, {0, 1, 1}, {1, 1, 0},   // coarse A
                            {1, 0, 0}, {1, 0, 1}, {1, 1, 1}, {0, 0, 1}};  // coarse B
    for (int s = 0; s < 8; ++s) {
        const int ctype = s / 4;
        t.type[s] = spec[s][0];
        t.dx[s] = spec[s][1];
        t.dy[s] = spec[s][2];
        const double ox = 0.5 * t.dx[s], oy = 0.5 * t.dy[s];
        double vx[3] = {ox, t.type[s] == 0 ? ox + 0.5 : ox, ox + 0.5};
        double vy[3] = {oy, t.type[s] == 0 ? oy : oy + 0.5, oy + 0.5};
        for (int c = 0; c < 3; ++c) {
            double l[3];
            coarse_barycentrics(ctype, vx[c], vy[c], l);
            for (int a = 0; a < 3; ++a) t.M[s][a][c] = l[a];
        }
    }
    return t;
}

// moments slot <-> symmetric 3x3
TM_HD int moment_slot(int a, int b) {
    if (a == b) return a;
    const int s = a + b;
    return s == 1 ? 3 : (s == 3 ? 4 : 5);
}

}  // namespace tmx
