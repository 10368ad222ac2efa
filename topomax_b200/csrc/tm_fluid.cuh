// FEM fluid problem of the reference: Stokes-Brinkman state equation on
// Taylor-Hood P2/P1 elements of the same right-diagonal mesh, dissipated-power objective and its
// sensitivity (reference: FEM_src/fluid_problem.py:68-122, src/penalizers.py:49-68).
//
//   state      [ M_r + K   -D^T ] [u]   [ -(M_r + K) g ]     g = prescribed boundary velocities
//              [  -D         0  ] [p] = [  D g - mean   ]     (all boundary nodes are Dirichlet)
//   M_r = int r(rho_h) phi_i phi_j (r rational in rho_h: FIAT's 12-point degree-6 rule, as FFC picks),
//   K = int grad phi_i . grad phi_j,  D_kj = int lambda_k d_d phi_j.  For test functions that vanish
//   on the boundary the reference's (grad p, v) block equals -D^T, so the system is the symmetric
//   saddle point above; the constant-pressure null space is handled by projecting the continuity
//   right-hand side (the reference's own system is singular and, for its diffuser design,
//   inconsistent: see oracle/fluid_oracle.py).
//   objective  1/2 (u^T M_r u + mu u^T K u);  sensitivity rhs_i = int 1/2 r'(rho_h) |u|^2 lambda_i
//   (4x4 collapsed Gauss-Jacobi, degree 7), then the P1 mass solve of the L2 projection.
//
// FIRST CORRECT VERSION of this row: one thread per triangle, contributions scattered with fp64
// atomics (summation order not fixed), weighted mass matrices stored per solve (21 numbers per
// triangle), diagonally preconditioned MINRES with the Lanczos scalars read back every iteration.
// Element arithmetic and the MINRES loop are host-compilable and are checked on the CPU against the
// oracle (tests/hostcheck/fluid_host.cpp); the CUDA back-end only adds launches.
#pragma once

#include <cmath>
#include <cstddef>

#include "tm_element.cuh"  // TM_HD
#include "tm_trimg.cuh"    // node -> incident triangles (gather forms)

namespace tmx {

struct FluidTables {
    double area;
    double grad_lam[2][3][2];   // [triangle type A/B][barycentric][x|y]
    double Kref[2][6][6];       // int grad phi_i . grad phi_j
    double Dloc[2][3][6][2];    // int lambda_c d_d phi_j
    double phi12[12][6], lam12[12][3], w12[12];  // state form / objective rule (weights sum to 1)
    double phi16[16][6], lam16[16][3], w16[16];  // sensitivity rule
};

struct FluidGeom {
    int nx, ny, Lx, Ly;
    double rmin, rmax, q;  // r(rho) = rmax + (rmin - rmax) rho (1+q)/(rho+q)
    double viscosity;
};

// P2 basis on barycentrics; local nodes 0,1,2 vertices, 3 = mid(0,1), 4 = mid(1,2), 5 = mid(0,2)
inline void fluid_p2_basis(const double (&l)[3], double (&phi)[6]) {
    phi[0] = l[0] * (2 * l[0] - 1);
    phi[1] = l[1] * (2 * l[1] - 1);
    phi[2] = l[2] * (2 * l[2] - 1);
    phi[3] = 4 * l[0] * l[1];
    phi[4] = 4 * l[1] * l[2];
    phi[5] = 4 * l[0] * l[2];
}
inline void fluid_p2_grad(const double (&l)[3], const double (&gl)[3][2], double (&dphi)[6][2]) {
    const int ea[3] = {0, 1, 0}, eb[3] = {1, 2, 2};
    for (int d = 0; d < 2; ++d) {
        for (int a = 0; a < 3; ++a) dphi[a][d] = (4 * l[a] - 1) * gl[a][d];
        for (int k = 0; k < 3; ++k) dphi[3 + k][d] = 4 * (l[ea[k]] * gl[eb[k]][d] + l[eb[k]] * gl[ea[k]][d]);
    }
}

// FIAT's default schemes (create_quadrature): degree 6 -> 12 points; degree 7 -> collapsed
// Gauss-Jacobi with (7+2)/2 = 4 points per axis.  Barycentric (1-X-Y, X, Y), weights sum to 1.
inline void fluid_rule12(double (&lam)[12][3], double (&w)[12]) {
    const double a1 = 0.873821971016996, b1 = 0.063089014491502;
    const double a2 = 0.501426509658179, b2 = 0.249286745170910;
    const double c1 = 0.636502499121399, c2 = 0.310352451033785, c3 = 0.053145049844816;
    const double xy[12][2] = {{b1, b1}, {a1, b1}, {b1, a1}, {b2, b2}, {a2, b2}, {b2, a2},
                              {c2, c3}, {c3, c2}, {c1, c3}, {c3, c1}, {c1, c2}, {c2, c1}};
    const double ww[3] = {0.050844906370207, 0.116786275726379, 0.082851075618374};
    double sum = 0.0;
    for (int k = 0; k < 12; ++k) {
        lam[k][0] = 1.0 - xy[k][0] - xy[k][1];
        lam[k][1] = xy[k][0];
        lam[k][2] = xy[k][1];
        w[k] = ww[k < 3 ? 0 : (k < 6 ? 1 : 2)];
        sum += w[k];
    }
    if (std::fabs(sum - 1.0) < 1e-12)  // the published table sums to 1 up to its 15 digits: normalise
        for (int k = 0; k < 12; ++k) w[k] /= sum;
}
inline void fluid_rule16(double (&lam)[16][3], double (&w)[16]) {
    const double e1[4] = {-0.8611363115940526, -0.3399810435848563, 0.3399810435848563, 0.8611363115940526};
    const double w1[4] = {0.3478548451374538, 0.6521451548625462, 0.6521451548625462, 0.3478548451374538};
    const double e2[4] = {-0.8857916077709646, -0.44631397272375245, 0.16718086473783364, 0.7204802713124389};
    const double w2[4] = {0.5420276537259541, 0.8138582720410844, 0.5193901904329293, 0.12472388380003234};
    int k = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j, ++k) {
            const double xi1 = 0.5 * (1.0 + e1[i]) * (1.0 - e2[j]) - 1.0, xi2 = e2[j];
            const double x = 0.5 * (xi1 + 1.0), y = 0.5 * (xi2 + 1.0);
            lam[k][0] = 1.0 - x - y;
            lam[k][1] = x;
            lam[k][2] = y;
            w[k] = 0.25 * w1[i] * w2[j];
        }
}

// mesh of nx x ny cells of size hx x hy; T_A = (v0, v1, v3), T_B = (v0, v2, v3)
inline FluidTables fluid_make_tables(double hx, double hy) {
    FluidTables T;
    T.area = 0.5 * hx * hy;
    // vertices of the two triangle types in cell-local coordinates
    const double P[2][3][2] = {{{0, 0}, {hx, 0}, {hx, hy}}, {{0, 0}, {0, hy}, {hx, hy}}};
    for (int t = 0; t < 2; ++t) {
        const double j00 = P[t][1][0] - P[t][0][0], j01 = P[t][2][0] - P[t][0][0];
        const double j10 = P[t][1][1] - P[t][0][1], j11 = P[t][2][1] - P[t][0][1];
        const double det = j00 * j11 - j01 * j10;
        // rows of J^-1 are grad lambda_1, grad lambda_2
        const double g1[2] = {j11 / det, -j01 / det}, g2[2] = {-j10 / det, j00 / det};
        for (int d = 0; d < 2; ++d) {
            T.grad_lam[t][0][d] = -(g1[d] + g2[d]);
            T.grad_lam[t][1][d] = g1[d];
            T.grad_lam[t][2][d] = g2[d];
        }
    }
    fluid_rule12(T.lam12, T.w12);
    fluid_rule16(T.lam16, T.w16);
    for (int k = 0; k < 12; ++k) fluid_p2_basis(T.lam12[k], T.phi12[k]);
    for (int k = 0; k < 16; ++k) fluid_p2_basis(T.lam16[k], T.phi16[k]);
    for (int t = 0; t < 2; ++t) {
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 6; ++j) T.Kref[t][i][j] = 0.0;
        for (int c = 0; c < 3; ++c)
            for (int j = 0; j < 6; ++j) T.Dloc[t][c][j][0] = T.Dloc[t][c][j][1] = 0.0;
        for (int k = 0; k < 12; ++k) {
            double dphi[6][2];
            fluid_p2_grad(T.lam12[k], T.grad_lam[t], dphi);
            const double wk = T.w12[k] * T.area;
            for (int i = 0; i < 6; ++i)
                for (int j = 0; j < 6; ++j)
                    T.Kref[t][i][j] += wk * (dphi[i][0] * dphi[j][0] + dphi[i][1] * dphi[j][1]);
            for (int c = 0; c < 3; ++c)
                for (int j = 0; j < 6; ++j)
                    for (int d = 0; d < 2; ++d) T.Dloc[t][c][j][d] += wk * T.lam12[k][c] * dphi[j][d];
        }
    }
    return T;
}

// ---- indexing: P2 lattice node = j*Lx + i, vertex = iy*(nx+1) + ix
TM_HD void fluid_tri_nodes(const FluidGeom& g, int cx, int cy, int t, int (&node)[6], int (&vert)[3],
                           bool (&interior)[6]) {
    // lattice offsets of local nodes 0..5 for T_A and T_B
    const int di[2][6] = {{0, 2, 2, 1, 2, 1}, {0, 0, 2, 0, 1, 1}};
    const int dj[2][6] = {{0, 0, 2, 0, 1, 1}, {0, 2, 2, 1, 2, 1}};
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const int i = 2 * cx + di[t][k], j = 2 * cy + dj[t][k];
        node[k] = j * g.Lx + i;
        interior[k] = i > 0 && j > 0 && i < g.Lx - 1 && j < g.Ly - 1;
    }
    const int v0 = cy * (g.nx + 1) + cx;
    vert[0] = v0;
    vert[1] = t == 0 ? v0 + 1 : v0 + g.nx + 1;
    vert[2] = v0 + g.nx + 2;
}

TM_HD double fluid_r(const FluidGeom& g, double rho) {
    return g.rmax + (g.rmin - g.rmax) * rho * (1 + g.q) / (rho + g.q);
}
TM_HD double fluid_r_prime(const FluidGeom& g, double rho) {
    return (g.rmin - g.rmax) * g.q * (1 + g.q) / ((rho + g.q) * (rho + g.q));
}

TM_HD int fluid_sym(int i, int j) {  // packed upper triangle of a 6x6 symmetric matrix
    const int a = i < j ? i : j, b = i < j ? j : i;
    return a * 6 - a * (a - 1) / 2 + (b - a);
}

// weighted mass matrix of one triangle: Me[ij] = |T| sum_q w_q r(rho_h(q)) phi_i(q) phi_j(q)
TM_HD void fluid_tri_mass(const FluidTables& T, const FluidGeom& g, const double (&rho3)[3], double (&Me)[21]) {
#pragma unroll
    for (int k = 0; k < 21; ++k) Me[k] = 0.0;
    for (int q = 0; q < 12; ++q) {
        const double rho = T.lam12[q][0] * rho3[0] + T.lam12[q][1] * rho3[1] + T.lam12[q][2] * rho3[2];
        const double wr = T.w12[q] * T.area * fluid_r(g, rho);
        int k = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = i; j < 6; ++j, ++k) Me[k] += wr * T.phi12[q][i] * T.phi12[q][j];
    }
}

// local product of the symmetric saddle-point operator on one triangle
TM_HD void fluid_tri_apply(const FluidTables& T, int t, const double (&Me)[21], const double (&ux)[6],
                           const double (&uy)[6], const double (&p)[3], double (&yx)[6], double (&yy)[6],
                           double (&yp)[3]) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double sx = 0.0, sy = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const double a = Me[fluid_sym(i, j)] + T.Kref[t][i][j];
            sx += a * ux[j];
            sy += a * uy[j];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {  // -D^T p
            sx -= T.Dloc[t][c][i][0] * p[c];
            sy -= T.Dloc[t][c][i][1] * p[c];
        }
        yx[i] = sx;
        yy[i] = sy;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {  // -D u
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) s += T.Dloc[t][c][j][0] * ux[j] + T.Dloc[t][c][j][1] * uy[j];
        yp[c] = -s;
    }
}

// 1/2 (u^T M_r u + mu u^T K u) of one triangle
TM_HD double fluid_tri_objective(const FluidTables& T, int t, const double (&Me)[21], const double (&ux)[6],
                                 const double (&uy)[6], double viscosity) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const double a = Me[fluid_sym(i, j)] + viscosity * T.Kref[t][i][j];
            s += a * (ux[i] * ux[j] + uy[i] * uy[j]);
        }
    return 0.5 * s;
}

// out[c] = int_T 1/2 r'(rho_h) |u_h|^2 lambda_c   (16-point rule)
TM_HD void fluid_tri_sens(const FluidTables& T, const FluidGeom& g, const double (&rho3)[3],
                          const double (&ux)[6], const double (&uy)[6], double (&out)[3]) {
    out[0] = out[1] = out[2] = 0.0;
    for (int q = 0; q < 16; ++q) {
        const double rho = T.lam16[q][0] * rho3[0] + T.lam16[q][1] * rho3[1] + T.lam16[q][2] * rho3[2];
        double vx = 0.0, vy = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            vx += T.phi16[q][k] * ux[k];
            vy += T.phi16[q][k] * uy[k];
        }
        const double f = 0.5 * fluid_r_prime(g, rho) * (vx * vx + vy * vy) * T.w16[q] * T.area;
#pragma unroll
        for (int c = 0; c < 3; ++c) out[c] += f * T.lam16[q][c];
    }
}

// ---------------------------------------------------------------------------------------------
// Per-triangle work items (tid = 2*cell + type, cell = cy*nx + cx): gather, element arithmetic,
// scatter through `add` (atomicAdd on the device, += in the serial host check).  Vectors are the
// combined [u (2 per lattice node, interleaved) | p (per vertex)]; Me_all is [21][ntri].
// ---------------------------------------------------------------------------------------------
struct FluidSerialAdd {
    TM_HD void operator()(double* p, double v) const { *p += v; }
};

TM_HD void fluid_tid_to_cell(const FluidGeom& g, size_t tid, int& cx, int& cy, int& t) {
    t = (int)(tid & 1);
    const size_t cell = tid >> 1;
    cy = (int)(cell / g.nx);
    cx = (int)(cell - (size_t)cy * g.nx);
}

TM_HD void fluid_body_mass(const FluidTables& T, const FluidGeom& g, const double* rho, double* Me_all,
                           size_t ntri, size_t tid) {
    int cx, cy, t, node[6], vert[3];
    bool interior[6];
    fluid_tid_to_cell(g, tid, cx, cy, t);
    fluid_tri_nodes(g, cx, cy, t, node, vert, interior);
    const double rho3[3] = {rho[vert[0]], rho[vert[1]], rho[vert[2]]};
    double Me[21];
    fluid_tri_mass(T, g, rho3, Me);
#pragma unroll
    for (int k = 0; k < 21; ++k) Me_all[(size_t)k * ntri + tid] = Me[k];
}

// mode 0: y += Op x with the boundary velocities of x taken as zero and no boundary rows written
//         (boundary rows of the operator are the identity: the caller keeps them zero);
// mode 1: lifting, y += Op [g on the boundary, 0 inside | 0]: interior momentum rows and all
//         continuity rows of the prescribed boundary values in x.
template <class Add>
TM_HD void fluid_body_apply(const FluidTables& T, const FluidGeom& g, const double* Me_all, size_t ntri,
                            const double* x, double* y, size_t tid, int mode, Add add) {
    int cx, cy, t, node[6], vert[3];
    bool interior[6];
    fluid_tid_to_cell(g, tid, cx, cy, t);
    fluid_tri_nodes(g, cx, cy, t, node, vert, interior);
    const size_t nu = (size_t)2 * g.Lx * g.Ly;
    double Me[21], ux[6], uy[6], p[3], yx[6], yy[6], yp[3];
#pragma unroll
    for (int k = 0; k < 21; ++k) Me[k] = Me_all[(size_t)k * ntri + tid];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const bool use = mode == 0 ? interior[k] : !interior[k];
        ux[k] = use ? x[2 * (size_t)node[k]] : 0.0;
        uy[k] = use ? x[2 * (size_t)node[k] + 1] : 0.0;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) p[c] = mode == 0 ? x[nu + vert[c]] : 0.0;
    fluid_tri_apply(T, t, Me, ux, uy, p, yx, yy, yp);
#pragma unroll
    for (int k = 0; k < 6; ++k)
        if (interior[k]) {
            add(&y[2 * (size_t)node[k]], yx[k]);
            add(&y[2 * (size_t)node[k] + 1], yy[k]);
        }
#pragma unroll
    for (int c = 0; c < 3; ++c) add(&y[nu + vert[c]], yp[c]);
}

// diagonal of the velocity block (interior nodes; the caller presets 1 on boundary nodes)
template <class Add>
TM_HD void fluid_body_diag(const FluidTables& T, const FluidGeom& g, const double* Me_all, size_t ntri,
                           double* diag, size_t tid, Add add) {
    int cx, cy, t, node[6], vert[3];
    bool interior[6];
    fluid_tid_to_cell(g, tid, cx, cy, t);
    fluid_tri_nodes(g, cx, cy, t, node, vert, interior);
#pragma unroll
    for (int k = 0; k < 6; ++k)
        if (interior[k]) {
            const double a = Me_all[(size_t)fluid_sym(k, k) * ntri + tid] + T.Kref[t][k][k];
            add(&diag[2 * (size_t)node[k]], a);
            add(&diag[2 * (size_t)node[k] + 1], a);
        }
}

// pressure part of the preconditioner: sum over triangles of Dloc^2 / diag(A) (an un-assembled
// stand-in for diag(B diag(A)^-1 B^T); any positive diagonal keeps MINRES valid)
template <class Add>
TM_HD void fluid_body_schur(const FluidTables& T, const FluidGeom& g, double* diag, size_t tid, Add add) {
    int cx, cy, t, node[6], vert[3];
    bool interior[6];
    fluid_tid_to_cell(g, tid, cx, cy, t);
    fluid_tri_nodes(g, cx, cy, t, node, vert, interior);
    const size_t nu = (size_t)2 * g.Lx * g.Ly;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k)
            if (interior[k]) {
                s += T.Dloc[t][c][k][0] * T.Dloc[t][c][k][0] / diag[2 * (size_t)node[k]];
                s += T.Dloc[t][c][k][1] * T.Dloc[t][c][k][1] / diag[2 * (size_t)node[k] + 1];
            }
        add(&diag[nu + vert[c]], s);
    }
}

// local 3x3 matrix of the P1 "Darcy" Laplacian int (1/r) grad lambda_a . grad lambda_b (the pressure
// Schur complement where the Brinkman term dominates), 1/r averaged over the vertices, plus a
// relative 1e-8 on the diagonal so that the pure-Neumann operator is definite.  L_all is [6][ntri].
TM_HD void fluid_body_darcy(const FluidTables& T, const FluidGeom& g, const double* rho, double* L_all,
                            size_t ntri, size_t tid) {
    int cx, cy, t, node[6], vert[3];
    bool interior[6];
    fluid_tid_to_cell(g, tid, cx, cy, t);
    fluid_tri_nodes(g, cx, cy, t, node, vert, interior);
    const double w = T.area * (1.0 / fluid_r(g, rho[vert[0]]) + 1.0 / fluid_r(g, rho[vert[1]]) +
                               1.0 / fluid_r(g, rho[vert[2]])) / 3.0;
    double L[3][3], tr = 0.0;
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
            L[a][b] = w * (T.grad_lam[t][a][0] * T.grad_lam[t][b][0] + T.grad_lam[t][a][1] * T.grad_lam[t][b][1]);
    for (int a = 0; a < 3; ++a) tr += L[a][a];
    int k = 0;
    for (int a = 0; a < 3; ++a)
        for (int b = a; b < 3; ++b, ++k) L_all[(size_t)k * ntri + tid] = L[a][b] + (a == b ? 1e-8 * tr / 3.0 : 0.0);
}

// diagonal of the P1 mass matrix: |T|/6 per triangle and vertex
template <class Add>
TM_HD void fluid_body_pmass_diag(const FluidTables& T, const FluidGeom& g, double* diag, size_t tid, Add add) {
    int cx, cy, t, node[6], vert[3];
    bool interior[6];
    fluid_tid_to_cell(g, tid, cx, cy, t);
    fluid_tri_nodes(g, cx, cy, t, node, vert, interior);
    for (int c = 0; c < 3; ++c) add(&diag[vert[c]], T.area / 6.0);
}

TM_HD double fluid_body_objective(const FluidTables& T, const FluidGeom& g, const double* Me_all, size_t ntri,
                                  const double* u, size_t tid) {
    int cx, cy, t, node[6], vert[3];
    bool interior[6];
    fluid_tid_to_cell(g, tid, cx, cy, t);
    fluid_tri_nodes(g, cx, cy, t, node, vert, interior);
    double Me[21], ux[6], uy[6];
#pragma unroll
    for (int k = 0; k < 21; ++k) Me[k] = Me_all[(size_t)k * ntri + tid];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        ux[k] = u[2 * (size_t)node[k]];
        uy[k] = u[2 * (size_t)node[k] + 1];
    }
    return fluid_tri_objective(T, t, Me, ux, uy, g.viscosity);
}

template <class Add>
TM_HD void fluid_body_sens(const FluidTables& T, const FluidGeom& g, const double* rho, const double* u,
                           double* out, size_t tid, Add add) {
    int cx, cy, t, node[6], vert[3];
    bool interior[6];
    fluid_tid_to_cell(g, tid, cx, cy, t);
    fluid_tri_nodes(g, cx, cy, t, node, vert, interior);
    const double rho3[3] = {rho[vert[0]], rho[vert[1]], rho[vert[2]]};
    double ux[6], uy[6], o[3];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        ux[k] = u[2 * (size_t)node[k]];
        uy[k] = u[2 * (size_t)node[k] + 1];
    }
    fluid_tri_sens(T, g, rho3, ux, uy, o);
#pragma unroll
    for (int c = 0; c < 3; ++c) add(&out[vert[c]], o[c]);
}


// ---------------------------------------------------------------------------------------------
// Gather forms of the scatter work items above (opt-in "deterministic" mode): one work item per
// OUTPUT entry, which walks the triangles incident to its node (trimg_incident) in a fixed order and
// writes once -- no atomics, no memset, bit-reproducible sums.  Work item k < n2 is lattice node k
// (both velocity components), k >= n2 is vertex k - n2 (pressure / continuity row).
// ---------------------------------------------------------------------------------------------
TM_HD void fluid_body_apply_gather(const FluidTables& T, const FluidGeom& g, const double* Me_all, size_t ntri,
                                   const double* x, double* y, size_t k, int mode) {
    const size_t n2 = (size_t)g.Lx * g.Ly, nu = 2 * n2;
    size_t tri[6];
    int local[6];
    if (k < n2) {
        const TriLevel lv{g.nx, g.ny, 2, 1};
        const int cnt = trimg_incident<6>(lv, k, tri, local);
        double sx = 0.0, sy = 0.0;
        bool boundary = false;
        for (int q = 0; q < cnt; ++q) {
            int cx, cy, t, node[6], vert[3];
            bool interior[6];
            fluid_tid_to_cell(g, tri[q], cx, cy, t);
            fluid_tri_nodes(g, cx, cy, t, node, vert, interior);
            const int i = local[q];
            if (!interior[i]) {
                boundary = true;
                break;
            }
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const bool use = mode == 0 ? interior[j] : !interior[j];
                if (!use) continue;
                const double a = Me_all[(size_t)fluid_sym(i, j) * ntri + tri[q]] + T.Kref[t][i][j];
                sx += a * x[2 * (size_t)node[j]];
                sy += a * x[2 * (size_t)node[j] + 1];
            }
            if (mode == 0)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const double p = x[nu + vert[c]];
                    sx -= T.Dloc[t][c][i][0] * p;
                    sy -= T.Dloc[t][c][i][1] * p;
                }
        }
        y[2 * k] = boundary ? 0.0 : sx;
        y[2 * k + 1] = boundary ? 0.0 : sy;
    } else {
        const size_t v = k - n2;
        const TriLevel lv{g.nx, g.ny, 1, 0};
        const int cnt = trimg_incident<3>(lv, v, tri, local);
        double s = 0.0;
        for (int q = 0; q < cnt; ++q) {
            int cx, cy, t, node[6], vert[3];
            bool interior[6];
            fluid_tid_to_cell(g, tri[q], cx, cy, t);
            fluid_tri_nodes(g, cx, cy, t, node, vert, interior);
            const int c = local[q];
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const bool use = mode == 0 ? interior[j] : !interior[j];
                if (!use) continue;
                s += T.Dloc[t][c][j][0] * x[2 * (size_t)node[j]] + T.Dloc[t][c][j][1] * x[2 * (size_t)node[j] + 1];
            }
        }
        y[nu + v] = -s;
    }
}

// the whole preconditioner diagonal in two gather passes: velocity entries (pass 0), then the
// pressure entries, which read the velocity ones (pass 1)
TM_HD void fluid_body_diag_gather(const FluidTables& T, const FluidGeom& g, const double* Me_all, size_t ntri,
                                  double* diag, size_t k, int pass) {
    const size_t n2 = (size_t)g.Lx * g.Ly, nu = 2 * n2;
    size_t tri[6];
    int local[6];
    if (pass == 0) {
        if (k >= n2) return;
        const TriLevel lv{g.nx, g.ny, 2, 1};
        const int cnt = trimg_incident<6>(lv, k, tri, local);
        double a = 0.0;
        bool boundary = false;
        for (int q = 0; q < cnt; ++q) {
            int cx, cy, t, node[6], vert[3];
            bool interior[6];
            fluid_tid_to_cell(g, tri[q], cx, cy, t);
            fluid_tri_nodes(g, cx, cy, t, node, vert, interior);
            if (!interior[local[q]]) {
                boundary = true;
                break;
            }
            a += Me_all[(size_t)fluid_sym(local[q], local[q]) * ntri + tri[q]] + T.Kref[t][local[q]][local[q]];
        }
        diag[2 * k] = diag[2 * k + 1] = boundary ? 1.0 : a;
    } else {
        if (k < n2) return;
        const size_t v = k - n2;
        const TriLevel lv{g.nx, g.ny, 1, 0};
        const int cnt = trimg_incident<3>(lv, v, tri, local);
        double s = 0.0;
        for (int q = 0; q < cnt; ++q) {
            int cx, cy, t, node[6], vert[3];
            bool interior[6];
            fluid_tid_to_cell(g, tri[q], cx, cy, t);
            fluid_tri_nodes(g, cx, cy, t, node, vert, interior);
            const int c = local[q];
#pragma unroll
            for (int j = 0; j < 6; ++j)
                if (interior[j]) {
                    s += T.Dloc[t][c][j][0] * T.Dloc[t][c][j][0] / diag[2 * (size_t)node[j]];
                    s += T.Dloc[t][c][j][1] * T.Dloc[t][c][j][1] / diag[2 * (size_t)node[j] + 1];
                }
        }
        diag[nu + v] = s;
    }
}

// sensitivity right-hand side of vertex v, and the P1 mass diagonal, by gather
TM_HD void fluid_body_sens_gather(const FluidTables& T, const FluidGeom& g, const double* rho, const double* u,
                                  double* out, size_t v) {
    size_t tri[6];
    int local[6];
    const TriLevel lv{g.nx, g.ny, 1, 0};
    const int cnt = trimg_incident<3>(lv, v, tri, local);
    double s = 0.0;
    for (int q = 0; q < cnt; ++q) {
        int cx, cy, t, node[6], vert[3];
        bool interior[6];
        fluid_tid_to_cell(g, tri[q], cx, cy, t);
        fluid_tri_nodes(g, cx, cy, t, node, vert, interior);
        const double rho3[3] = {rho[vert[0]], rho[vert[1]], rho[vert[2]]};
        double ux[6], uy[6], o[3];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            ux[j] = u[2 * (size_t)node[j]];
            uy[j] = u[2 * (size_t)node[j] + 1];
        }
        fluid_tri_sens(T, g, rho3, ux, uy, o);
        s += o[local[q]];
    }
    out[v] = s;
}
TM_HD void fluid_body_pmass_diag_gather(const FluidTables& T, const FluidGeom& g, double* diag, size_t v) {
    size_t tri[6];
    int local[6];
    const TriLevel lv{g.nx, g.ny, 1, 0};
    diag[v] = trimg_incident<3>(lv, v, tri, local) * (T.area / 6.0);
}

// ---------------------------------------------------------------------------------------------
// Preconditioned MINRES (Elman, Silvester & Wathen, Alg. 4.1) on the combined vector [u | p].
// BK supplies: n(), apply(x, y), precond(r, z), dot(a, b), copy, zero, and the two fused updates;
// the same loop runs on the CUDA back-end and on the host back-end of the CPU check.
// ---------------------------------------------------------------------------------------------
struct MinresResult {
    int iterations = 0;
    double relres = 0.0;  // |eta_k| / |eta_0|, the preconditioned residual norm MINRES minimises
    bool converged = false;
};

// `ref_norm` > 0: the tolerance is relative to it instead of the initial residual norm (warm starts:
// b is then the residual of the initial guess and ref_norm the norm of the original right-hand side)
template <class BK>
MinresResult fluid_minres(BK& bk, const typename BK::Vec& b, typename BK::Vec& x, double rtol, int maxit,
                          double ref_norm = 0.0) {
    using Vec = typename BK::Vec;
    MinresResult res;
    Vec& v_old = bk.work(0);
    Vec& v = bk.work(1);
    Vec& v_new = bk.work(2);
    Vec& z = bk.work(3);
    Vec& z_new = bk.work(4);
    Vec& w_old = bk.work(5);
    Vec& w = bk.work(6);
    Vec& w_new = bk.work(7);
    bk.zero(x);
    bk.zero(v_old);
    bk.zero(w_old);
    bk.zero(w);
    bk.copy(b, v);
    bk.precond(v, z);
    double gamma = std::sqrt(bk.dot(z, v)), gamma_old = 1.0;
    if (!(gamma > 0.0)) {
        res.converged = true;
        return res;
    }
    double eta = gamma;
    const double eta0 = ref_norm > 0.0 ? ref_norm : gamma;
    if (gamma <= rtol * eta0) {  // the initial guess is already good enough
        res.relres = gamma / eta0;
        res.converged = true;
        return res;
    }
    double s_old = 0.0, s = 0.0, c_old = 1.0, c = 1.0;
    Vec* pv_old = &v_old;
    Vec* pv = &v;
    Vec* pv_new = &v_new;
    Vec* pz = &z;
    Vec* pz_new = &z_new;
    Vec* pw_old = &w_old;
    Vec* pw = &w;
    Vec* pw_new = &w_new;
    for (int j = 1; j <= maxit; ++j) {
        bk.scale(*pz, 1.0 / gamma);         // z_j /= gamma_j
        bk.apply(*pz, *pv_new);             // v_new = A z_j
        const double delta = bk.dot(*pv_new, *pz);
        // v_new -= (delta/gamma) v + (gamma/gamma_old) v_old
        bk.axpy2(*pv_new, -delta / gamma, *pv, -gamma / gamma_old, *pv_old);
        bk.precond(*pv_new, *pz_new);
        const double gamma_new = std::sqrt(std::fmax(bk.dot(*pz_new, *pv_new), 0.0));
        const double a0 = c * delta - c_old * s * gamma;
        const double a1 = std::sqrt(a0 * a0 + gamma_new * gamma_new);
        const double a2 = s * delta + c_old * c * gamma;
        const double a3 = s_old * gamma;
        const double c_new = a0 / a1, s_new = gamma_new / a1;
        // w_new = (z_j - a3 w_old - a2 w) / a1 ;  x += c_new eta w_new
        bk.direction(*pw_new, *pz, a3, *pw_old, a2, *pw, 1.0 / a1, x, c_new * eta);
        eta = -s_new * eta;
        res.iterations = j;
        res.relres = std::fabs(eta) / eta0;
        if (!(res.relres == res.relres)) break;  // NaN
        if (res.relres <= rtol) {
            res.converged = true;
            break;
        }
        // rotate
        Vec* t = pv_old; pv_old = pv; pv = pv_new; pv_new = t;
        t = pz; pz = pz_new; pz_new = t;
        t = pw_old; pw_old = pw; pw = pw_new; pw_new = t;
        gamma_old = gamma;
        gamma = gamma_new;
        c_old = c; c = c_new;
        s_old = s; s = s_new;
        if (!(gamma > 0.0)) {  // exact breakdown: the Krylov space is exhausted
            res.converged = true;
            break;
        }
    }
    return res;
}


// ---------------------------------------------------------------------------------------------
// The same MINRES with its Lanczos / Givens scalars RESIDENT ON THE DEVICE: the vector kernels read
// their coefficients from a small array S, two one-thread "scalar steps" advance the recurrences, and
// the host reads the residual estimate back only every `check_every` iterations (no synchronisation
// per iteration: the prerequisite for replaying an iteration from a CUDA graph).  Iterations past
// convergence inside a check interval are harmless (the residual norm is non-increasing).
// ---------------------------------------------------------------------------------------------
enum {
    MS_GAMMA = 0, MS_GAMMA_OLD = 1, MS_ETA = 2, MS_ETA0 = 3, MS_S_OLD = 4, MS_S = 5, MS_C_OLD = 6, MS_C = 7,
    MS_DELTA = 8, MS_GNEW_SQ = 9, MS_COEF_V = 10, MS_COEF_VOLD = 11, MS_A3 = 12, MS_A2 = 13, MS_INV_A1 = 14,
    MS_STEP = 15, MS_INV_GAMMA = 16, MS_RELRES = 17, MS_BREAKDOWN = 18, MS_ITER = 19, MS_COUNT = 24
};

// start: S[MS_GNEW_SQ] holds z1 . v1
TM_HD void minres_scalar_init(double* S, double ref_norm) {
    const double gamma = sqrt(S[MS_GNEW_SQ] > 0.0 ? S[MS_GNEW_SQ] : 0.0);
    S[MS_GAMMA] = gamma;
    S[MS_GAMMA_OLD] = 1.0;
    S[MS_ETA] = gamma;
    S[MS_ETA0] = ref_norm > 0.0 ? ref_norm : gamma;
    S[MS_S_OLD] = 0.0;
    S[MS_S] = 0.0;
    S[MS_C_OLD] = 1.0;
    S[MS_C] = 1.0;
    S[MS_INV_GAMMA] = gamma > 0.0 ? 1.0 / gamma : 0.0;
    S[MS_RELRES] = S[MS_ETA0] > 0.0 ? gamma / S[MS_ETA0] : 0.0;
    S[MS_BREAKDOWN] = gamma > 0.0 ? 0.0 : 1.0;
    S[MS_ITER] = 0.0;
}
// after delta = (A z) . z: coefficients of  v_new -= (delta/gamma) v + (gamma/gamma_old) v_old
TM_HD void minres_scalar_a(double* S) {
    const double gamma = S[MS_GAMMA], gamma_old = S[MS_GAMMA_OLD];
    S[MS_COEF_V] = gamma > 0.0 ? -S[MS_DELTA] / gamma : 0.0;
    S[MS_COEF_VOLD] = gamma_old > 0.0 ? -gamma / gamma_old : 0.0;
}
// after gamma_new^2 = z_new . v_new: Givens rotation, direction coefficients, residual, rotation of names
TM_HD void minres_scalar_b(double* S) {
    const double gamma = S[MS_GAMMA], delta = S[MS_DELTA];
    const double gamma_new = sqrt(S[MS_GNEW_SQ] > 0.0 ? S[MS_GNEW_SQ] : 0.0);
    const double c = S[MS_C], c_old = S[MS_C_OLD], s = S[MS_S], s_old = S[MS_S_OLD];
    const double a0 = c * delta - c_old * s * gamma;
    const double a1 = sqrt(a0 * a0 + gamma_new * gamma_new);
    const double a2 = s * delta + c_old * c * gamma;
    const double a3 = s_old * gamma;
    const bool dead = S[MS_BREAKDOWN] != 0.0 || !(a1 > 0.0);
    const double c_new = dead ? 1.0 : a0 / a1, s_new = dead ? 0.0 : gamma_new / a1;
    S[MS_A3] = a3;
    S[MS_A2] = a2;
    S[MS_INV_A1] = dead ? 0.0 : 1.0 / a1;
    S[MS_STEP] = dead ? 0.0 : c_new * S[MS_ETA];
    if (!dead) {
        S[MS_ETA] = -s_new * S[MS_ETA];
        S[MS_ITER] += 1.0;
    }
    S[MS_GAMMA_OLD] = gamma;
    S[MS_GAMMA] = gamma_new;
    S[MS_C_OLD] = c;
    S[MS_C] = c_new;
    S[MS_S_OLD] = s;
    S[MS_S] = s_new;
    S[MS_INV_GAMMA] = gamma_new > 0.0 ? 1.0 / gamma_new : 0.0;
    if (!(gamma_new > 0.0)) S[MS_BREAKDOWN] = 1.0;  // Krylov space exhausted: later iterations are no-ops
    const double eta = S[MS_ETA] < 0.0 ? -S[MS_ETA] : S[MS_ETA];
    S[MS_RELRES] = S[MS_ETA0] > 0.0 ? eta / S[MS_ETA0] : 0.0;
}

// BK supplies the *_dev operations (coefficients named by their index in the device array) and
// read_scalars(host double[MS_COUNT]).
template <class BK>
MinresResult fluid_minres_dev(BK& bk, const typename BK::Vec& b, typename BK::Vec& x, double rtol, int maxit,
                              int check_every, double ref_norm = 0.0) {
    using Vec = typename BK::Vec;
    MinresResult res;
    Vec* pv_old = &bk.work(0);
    Vec* pv = &bk.work(1);
    Vec* pv_new = &bk.work(2);
    Vec* pz = &bk.work(3);
    Vec* pz_new = &bk.work(4);
    Vec* pw_old = &bk.work(5);
    Vec* pw = &bk.work(6);
    Vec* pw_new = &bk.work(7);
    bk.zero(x);
    bk.zero(*pv_old);
    bk.zero(*pw_old);
    bk.zero(*pw);
    bk.copy(b, *pv);
    bk.precond(*pv, *pz);
    bk.dot_dev(*pz, *pv, MS_GNEW_SQ);
    bk.scalar_init(ref_norm);
    double S[MS_COUNT];
    bk.read_scalars(S);
    res.relres = S[MS_RELRES];
    if (S[MS_BREAKDOWN] != 0.0 || S[MS_RELRES] <= rtol) {
        res.converged = true;
        return res;
    }
    if (check_every < 1) check_every = 1;
    for (int j = 1; j <= maxit; ++j) {
        bk.scale_dev(*pz, MS_INV_GAMMA);
        bk.apply(*pz, *pv_new);
        bk.dot_dev(*pv_new, *pz, MS_DELTA);
        bk.scalar_a();
        bk.axpy2_dev(*pv_new, MS_COEF_V, *pv, MS_COEF_VOLD, *pv_old);
        bk.precond(*pv_new, *pz_new);
        bk.dot_dev(*pz_new, *pv_new, MS_GNEW_SQ);
        bk.scalar_b();
        bk.direction_dev(*pw_new, *pz, MS_A3, *pw_old, MS_A2, *pw, MS_INV_A1, x, MS_STEP);
        Vec* t = pv_old; pv_old = pv; pv = pv_new; pv_new = t;
        t = pz; pz = pz_new; pz_new = t;
        t = pw_old; pw_old = pw; pw = pw_new; pw_new = t;
        if (j % check_every == 0 || j == maxit) {
            bk.read_scalars(S);
            res.iterations = (int)S[MS_ITER];
            res.relres = S[MS_RELRES];
            if (!(res.relres == res.relres)) break;
            if (res.relres <= rtol || S[MS_BREAKDOWN] != 0.0) {
                res.converged = true;
                break;
            }
        }
    }
    return res;
}


// ---------------------------------------------------------------------------------------------
// ... and replayed from a captured graph.  The vector roles rotate with periods 3 (v), 2 (z) and 3 (w),
// so after SIX iterations every pointer is back where it started: six iterations are captured once
// per solve (bk.begin_capture / end_capture: nothing executes while capturing) and the graph is
// replayed; the host looks at the residual every `check_every` iterations, rounded to whole blocks.
// ---------------------------------------------------------------------------------------------
constexpr int kMinresGraphBlock = 6;

template <class BK>
MinresResult fluid_minres_graph(BK& bk, const typename BK::Vec& b, typename BK::Vec& x, double rtol, int maxit,
                                int check_every, double ref_norm = 0.0) {
    using Vec = typename BK::Vec;
    MinresResult res;
    Vec* pv_old = &bk.work(0);
    Vec* pv = &bk.work(1);
    Vec* pv_new = &bk.work(2);
    Vec* pz = &bk.work(3);
    Vec* pz_new = &bk.work(4);
    Vec* pw_old = &bk.work(5);
    Vec* pw = &bk.work(6);
    Vec* pw_new = &bk.work(7);
    bk.zero(x);
    bk.zero(*pv_old);
    bk.zero(*pw_old);
    bk.zero(*pw);
    bk.copy(b, *pv);
    bk.precond(*pv, *pz);
    bk.dot_dev(*pz, *pv, MS_GNEW_SQ);
    bk.scalar_init(ref_norm);
    double S[MS_COUNT];
    bk.read_scalars(S);
    res.relres = S[MS_RELRES];
    if (S[MS_BREAKDOWN] != 0.0 || S[MS_RELRES] <= rtol) {
        res.converged = true;
        return res;
    }
    bk.begin_capture();
    for (int k = 0; k < kMinresGraphBlock; ++k) {
        bk.scale_dev(*pz, MS_INV_GAMMA);
        bk.apply(*pz, *pv_new);
        bk.dot_dev(*pv_new, *pz, MS_DELTA);
        bk.scalar_a();
        bk.axpy2_dev(*pv_new, MS_COEF_V, *pv, MS_COEF_VOLD, *pv_old);
        bk.precond(*pv_new, *pz_new);
        bk.dot_dev(*pz_new, *pv_new, MS_GNEW_SQ);
        bk.scalar_b();
        bk.direction_dev(*pw_new, *pz, MS_A3, *pw_old, MS_A2, *pw, MS_INV_A1, x, MS_STEP);
        Vec* t = pv_old; pv_old = pv; pv = pv_new; pv_new = t;
        t = pz; pz = pz_new; pz_new = t;
        t = pw_old; pw_old = pw; pw = pw_new; pw_new = t;
    }
    bk.end_capture();
    const int blocks_per_check = check_every > kMinresGraphBlock ? check_every / kMinresGraphBlock : 1;
    int blocks = 0;
    for (int it = 0; it < maxit; it += kMinresGraphBlock) {
        bk.replay();
        ++blocks;
        if (blocks % blocks_per_check == 0 || it + kMinresGraphBlock >= maxit) {
            bk.read_scalars(S);
            res.iterations = (int)S[MS_ITER];
            res.relres = S[MS_RELRES];
            if (!(res.relres == res.relres)) break;
            if (res.relres <= rtol || S[MS_BREAKDOWN] != 0.0) {
                res.converged = true;
                break;
            }
        }
    }
    bk.drop_graph();
    return res;
}

}  // namespace tmx
