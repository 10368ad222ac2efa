// Element-level arithmetic of the hot path, shared by every kernel (and compiled for the
// host by tests/hostcheck to be checked against the oracle without a GPU).
//
// Discretisation (reference: FEM_src/elasisity_problem.py:194-196, FEM_src/solver.py:38-49):
// vector-P2 displacement / nodal-P1 density on the "right"-diagonal structured triangle mesh.
// A cell owns the 3x3 block of the half-step lattice, local node q = 3*row + col:
//
//      6---7---8        T_A = (q0, q2, q8)  mids m01=q1  m12=q5  m02=q4   (lower right)
//      |  /    |        T_B = (q0, q6, q8)  mids m01=q3  m12=q7  m02=q4   (upper left)
//      3   4   5
//      | /     |
//      0---1---2
//
// State operator (reference: FEM_src/elasisity_problem.py:112-118, src/penalizers.py:36-40):
//   K_e = int_T r(xi_h) B^T D B,  r = m + (1-m) xi^3,  xi_h the P1 interpolant (varies inside T).
// grad(u_h) of a P2 field is affine in the barycentrics, grad u_h = sum_c lambda_c G_c, so
//   K_e = 2|T| sum_ab w_ab B_a^T D B_b ,   w_ab = 1/(2|T|) int_T r lambda_a lambda_b   (exact),
// six numbers per triangle.  The same six "moments" represent the EXACT Galerkin coarse
// operators of the multigrid hierarchy (tm_mg.cu), where they are stored instead of derived
// from xi.
#pragma once

#include <cmath>

#ifdef __CUDACC__
#define TM_HD __host__ __device__ __forceinline__
#else
#define TM_HD inline
#endif

namespace tmx {

// constitutive constants folded with the cell aspect ratio (square cells: kappa = 1,
// A11 = A22 = lambda + 2 mu, A12 = lambda, A33 = mu).  K_e is independent of the mesh size.
template <typename T>
struct Material {
    T A11, A22, A12, A33, kappa;
};

template <typename T>
inline Material<T> make_material(double lambda, double mu, double hx, double hy) {
    Material<T> m;
    m.A11 = T((lambda + 2 * mu) * hy / hx);
    m.A22 = T((lambda + 2 * mu) * hx / hy);
    m.A12 = T(lambda);
    m.A33 = T(mu * hy / hx);
    m.kappa = T(hx / hy);
    return m;
}

// moment slots: 0:w00 1:w11 2:w22 3:w01 4:w12 5:w02
// w_ab = c_ab [ m + (1-m) q_ab / 210 ],  c_ab = (1+delta_ab)/24,
// q_ab = p1 (p1^2 + 3 p2) + 2 p3,  p_k = sum_c xi_c^k + xi_a^k + xi_b^k
// (third moment of a linear form under Dirichlet(1+e_a+e_b); exact for the degree-5 integrand).
template <typename T>
TM_HD void moments_from_xi(T x0, T x1, T x2, T m, T w[6]) {
    const T a0 = x0 * x0, a1 = x1 * x1, a2 = x2 * x2;
    const T c0 = a0 * x0, c1 = a1 * x1, c2 = a2 * x2;
    const T s1 = x0 + x1 + x2, s2 = a0 + a1 + a2, s3 = c0 + c1 + c2;
    const T k = (T(1) - m) * T(1.0 / 210.0);
    const T d = T(1.0 / 12.0), o = T(1.0 / 24.0);
#define TM_Q(xa, xb, aa, ab, ca, cb, out, cc)                      \
    {                                                              \
        const T p1 = s1 + (xa) + (xb), p2 = s2 + (aa) + (ab);      \
        const T p3 = s3 + (ca) + (cb);                             \
        const T q = p1 * (p1 * p1 + T(3) * p2) + T(2) * p3;        \
        out = (cc) * (m + k * q);                                  \
    }
    TM_Q(x0, x0, a0, a0, c0, c0, w[0], d)
    TM_Q(x1, x1, a1, a1, c1, c1, w[1], d)
    TM_Q(x2, x2, a2, a2, c2, c2, w[2], d)
    TM_Q(x0, x1, a0, a1, c0, c1, w[3], o)
    TM_Q(x1, x2, a1, a2, c1, c2, w[4], o)
    TM_Q(x0, x2, a0, a2, c0, c2, w[5], o)
#undef TM_Q
}

// One triangle: adds K_e * (U, M) into (YU, YM).  U[c] vertex values, M[0..2] the edge
// midpoint values of edges (01), (12), (02); every value is a 2-vector.
// SWAP=false is T_A (d/dx along edge 0->1, d/dy along 1->2), SWAP=true is T_B (x <-> y).
template <typename T, bool SWAP>
TM_HD void tri_apply(const T U[3][2], const T M[3][2], const T w[6], const Material<T>& mat,
                     T YU[3][2], T YM[3][2]) {
    // C[c][b]: coefficient of grad(lambda_b) in grad u_h at vertex c
    T e[3][3];  // scaled strains at the three vertices: (g_xx, g_yy, gamma)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        T C[3][2];
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            if (b == c) {
                C[b][0] = T(3) * U[c][0];
                C[b][1] = T(3) * U[c][1];
            } else {
                const int mid = (b + c == 1) ? 0 : ((b + c == 3) ? 1 : 2);
                C[b][0] = T(4) * M[mid][0] - U[b][0];
                C[b][1] = T(4) * M[mid][1] - U[b][1];
            }
        }
        const T d1x = C[1][0] - C[0][0], d1y = C[1][1] - C[0][1];
        const T d2x = C[2][0] - C[1][0], d2y = C[2][1] - C[1][1];
        const T gxx = SWAP ? d2x : d1x, gxy = SWAP ? d2y : d1y;  // h_x d(u_x,u_y)/dx
        const T gyx = SWAP ? d1x : d2x, gyy = SWAP ? d1y : d2y;  // h_y d(u_x,u_y)/dy
        e[c][0] = gxx;
        e[c][1] = gyy;
        e[c][2] = gxy + mat.kappa * gyx;
    }
    T F[3][3][2];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const T wa0 = (a == 0) ? w[0] : ((a == 1) ? w[3] : w[5]);
        const T wa1 = (a == 0) ? w[3] : ((a == 1) ? w[1] : w[4]);
        const T wa2 = (a == 0) ? w[5] : ((a == 1) ? w[4] : w[2]);
        const T t0 = wa0 * e[0][0] + wa1 * e[1][0] + wa2 * e[2][0];
        const T t1 = wa0 * e[0][1] + wa1 * e[1][1] + wa2 * e[2][1];
        const T t2 = wa0 * e[0][2] + wa1 * e[1][2] + wa2 * e[2][2];
        // conjugates of (gx, gy): P = (Pxx, Pxy), Q = (Qyx, Qyy)
        const T Pxx = mat.A11 * t0 + mat.A12 * t1;
        const T Qyy = mat.A12 * t0 + mat.A22 * t1;
        const T Pxy = mat.A33 * t2;
        const T Qyx = mat.kappa * Pxy;
        const T S1x = SWAP ? Qyx : Pxx, S1y = SWAP ? Qyy : Pxy;  // conjugate of d1
        const T S2x = SWAP ? Pxx : Qyx, S2y = SWAP ? Pxy : Qyy;  // conjugate of d2
        F[a][0][0] = -S1x;       F[a][0][1] = -S1y;
        F[a][1][0] = S1x - S2x;  F[a][1][1] = S1y - S2y;
        F[a][2][0] = S2x;        F[a][2][1] = S2y;
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        YU[0][k] += T(3) * F[0][0][k] - F[1][0][k] - F[2][0][k];
        YU[1][k] += T(3) * F[1][1][k] - F[0][1][k] - F[2][1][k];
        YU[2][k] += T(3) * F[2][2][k] - F[0][2][k] - F[1][2][k];
        YM[0][k] += T(4) * (F[0][1][k] + F[1][0][k]);
        YM[1][k] += T(4) * (F[1][2][k] + F[2][1][k]);
        YM[2][k] += T(4) * (F[0][2][k] + F[2][0][k]);
    }
}

// Both triangles of a cell: acc[q] += (K_A + K_B) X, q the cell-local lattice node.
template <typename T>
TM_HD void cell_apply(const T X[9][2], const T wA[6], const T wB[6], const Material<T>& mat,
                      T acc[9][2]) {
    {
        const T U[3][2] = {{X[0][0], X[0][1]}, {X[2][0], X[2][1]}, {X[8][0], X[8][1]}};
        const T M[3][2] = {{X[1][0], X[1][1]}, {X[5][0], X[5][1]}, {X[4][0], X[4][1]}};
        T YU[3][2] = {{0, 0}, {0, 0}, {0, 0}}, YM[3][2] = {{0, 0}, {0, 0}, {0, 0}};
        tri_apply<T, false>(U, M, wA, mat, YU, YM);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            acc[0][k] += YU[0][k]; acc[2][k] += YU[1][k]; acc[8][k] += YU[2][k];
            acc[1][k] += YM[0][k]; acc[5][k] += YM[1][k]; acc[4][k] += YM[2][k];
        }
    }
    {
        const T U[3][2] = {{X[0][0], X[0][1]}, {X[6][0], X[6][1]}, {X[8][0], X[8][1]}};
        const T M[3][2] = {{X[3][0], X[3][1]}, {X[7][0], X[7][1]}, {X[4][0], X[4][1]}};
        T YU[3][2] = {{0, 0}, {0, 0}, {0, 0}}, YM[3][2] = {{0, 0}, {0, 0}, {0, 0}};
        tri_apply<T, true>(U, M, wB, mat, YU, YM);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            acc[0][k] += YU[0][k]; acc[6][k] += YU[1][k]; acc[8][k] += YU[2][k];
            acc[3][k] += YM[0][k]; acc[7][k] += YM[1][k]; acc[4][k] += YM[2][k];
        }
    }
}

// cell-local lattice node (0..8) of triangle-local node (0..5): vertices 0,1,2 then mids 01,12,02
TM_HD int tri_local_to_cell(int type, int k) {
    const int A[6] = {0, 2, 8, 1, 5, 4};
    const int B[6] = {0, 6, 8, 3, 7, 4};
    return type == 0 ? A[k] : B[k];
}
// inverse: triangle-local index of cell-local node q, or -1
TM_HD int cell_to_tri_local(int type, int q) {
    const int A[9] = {0, 3, 1, -1, 5, 4, -1, -1, 2};
    const int B[9] = {0, -1, -1, 3, 5, -1, 1, 4, 2};
    return type == 0 ? A[q] : B[q];
}

// Scaled strains (g_xx, g_yy, gamma) at the three vertices of one triangle, as in tri_apply.
template <typename T, bool SWAP>
TM_HD void tri_vertex_strains(const T U[3][2], const T M[3][2], T kappa, T e[3][3]) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        T C[3][2];
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            if (b == c) {
                C[b][0] = T(3) * U[c][0];
                C[b][1] = T(3) * U[c][1];
            } else {
                const int mid = (b + c == 1) ? 0 : ((b + c == 3) ? 1 : 2);
                C[b][0] = T(4) * M[mid][0] - U[b][0];
                C[b][1] = T(4) * M[mid][1] - U[b][1];
            }
        }
        const T d1x = C[1][0] - C[0][0], d1y = C[1][1] - C[0][1];
        const T d2x = C[2][0] - C[1][0], d2y = C[2][1] - C[1][1];
        const T gxx = SWAP ? d2x : d1x, gxy = SWAP ? d2y : d1y;
        const T gyx = SWAP ? d1x : d2x, gyy = SWAP ? d1y : d2y;
        e[c][0] = gxx;
        e[c][1] = gyy;
        e[c][2] = gxy + kappa * gyx;
    }
}

// Sensitivity right-hand side of one triangle (reference: FEM_src/elasisity_problem.py:146-150):
//   g_i = int_T -r'(xi_h) (lambda (div u)^2 + 2 mu eps:eps) phi_i^{P1},  r' = 3 (1-m) xi^2.
// With the energy density = sum_cd lambda_c lambda_d E_cd (E the strain energy form between
// vertex strains) and  1/(2|T|) int xi_h^2 lambda_c lambda_d lambda_i
//   = alpha!/120 * (p1^2 + p2)/42,   p_k = sum_v xi_v^k + xi_c^k + xi_d^k + xi_i^k.
template <typename T>
TM_HD void tri_sensitivity(const T e[3][3], const T xi[3], T m, const Material<T>& mat, T g[3]) {
    T E[3][3];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int d = c; d < 3; ++d) {
            E[c][d] = mat.A11 * e[c][0] * e[d][0] + mat.A22 * e[c][1] * e[d][1] +
                      mat.A12 * (e[c][0] * e[d][1] + e[c][1] * e[d][0]) + mat.A33 * e[c][2] * e[d][2];
            E[d][c] = E[c][d];
        }
    const T s1 = xi[0] + xi[1] + xi[2];
    const T s2 = xi[0] * xi[0] + xi[1] * xi[1] + xi[2] * xi[2];
    const T scale = -T(3) * (T(1) - m) / (T(120) * T(42));
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        T acc = 0;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                int cnt[3] = {0, 0, 0};
                cnt[c]++; cnt[d]++; cnt[i]++;
                const int fact[4] = {1, 1, 2, 6};
                const T af = T(fact[cnt[0]] * fact[cnt[1]] * fact[cnt[2]]);
                const T p1 = s1 + xi[c] + xi[d] + xi[i];
                const T p2 = s2 + xi[c] * xi[c] + xi[d] * xi[d] + xi[i] * xi[i];
                acc += E[c][d] * af * (p1 * p1 + p2);
            }
        g[i] = scale * acc;
    }
}

// ---------------------------------------------------------------------------------------
// General SIMP exponent (reference: src/penalizers.py:36-46, the penalties loop of
// src/solver.py:230-231).  The p = 3 closed forms above are the fast path of the fine-level
// operator; any other p goes through the functions below, evaluated ONCE per solve into stored
// level-0 moments (the operator then runs its stored-moment variant, like the coarse levels).
//
//  * integer p in [1, kMaxIntPenalty]: exact.  xi_h^p is a polynomial, and the p-th moment of a
//    linear form under a Dirichlet distribution with integer parameters is
//        E[(xi . lambda)^k] = k! / (n (n+1) .. (n+k-1)) * h_k(xs),
//    h_k the complete homogeneous symmetric polynomial of the n variables xs (vertex densities
//    repeated alpha_c times), from the power sums by Newton's identities.  (What FFC does for an
//    integer exponent: exact quadrature of degree p + 2.)
//  * any other p > 0: 16-point collapsed Gauss rule (degree 6), the same rule the oracle uses.
//    PARITY UNPINNED: the reference would integrate with FFC's default scheme for UFL's estimated
//    degree (heuristic degree(xi^p) = degree(xi) + 2); the integrand is not polynomial, so the
//    two differ at quadrature-error level.
// ---------------------------------------------------------------------------------------
constexpr int kMaxIntPenalty = 16;

struct PenaltySpec {
    double p;
    int ip;  // p when it is an integer in [1, kMaxIntPenalty], else 0 (quadrature)
};

TM_HD double dirichlet_power_moment(int k, const double* xs, int n) {
    double pw[6], ps[kMaxIntPenalty + 1], h[kMaxIntPenalty + 1];
    for (int j = 0; j < n; ++j) pw[j] = 1.0;
    h[0] = 1.0;
    double f = 1.0;
    for (int i = 1; i <= k; ++i) {
        double sum = 0.0;
        for (int j = 0; j < n; ++j) {
            pw[j] *= xs[j];
            sum += pw[j];
        }
        ps[i] = sum;
        double acc = 0.0;
        for (int t = 1; t <= i; ++t) acc += ps[t] * h[i - t];
        h[i] = acc / (double)i;
        f *= (double)i / (double)(n + i - 1);
    }
    return f * h[k];
}

// barycentric points / weights (sum 1) of the 4 x 4 collapsed Gauss rule: (x, y) = (a, b (1 - a))
TM_HD void collapsed_gauss16(int q, double lam[3], double& wt) {
    const double g[4] = {0.06943184420297371, 0.33000947820757187, 0.66999052179242813, 0.93056815579702629};
    const double w[4] = {0.17392742256872693, 0.32607257743127307, 0.32607257743127307, 0.17392742256872693};
    const double a = g[q >> 2], b = g[q & 3];
    const double x = a, y = b * (1.0 - a);
    lam[0] = 1.0 - x - y;
    lam[1] = x;
    lam[2] = y;
    wt = w[q >> 2] * w[q & 3] * (1.0 - a) * 2.0;
}

// w_ab = (2|T|)^-1 int r(xi_h) lambda_a lambda_b,  r = m + (1 - m) xi^p  (slots as moments_from_xi)
TM_HD void moments_general(double x0, double x1, double x2, double m, const PenaltySpec& ps, double w[6]) {
    const double x[3] = {x0, x1, x2};
    const int sa[6] = {0, 1, 2, 0, 1, 0}, sb[6] = {0, 1, 2, 1, 2, 2};
    if (ps.ip > 0) {
        for (int k = 0; k < 6; ++k) {
            const double xs[5] = {x0, x1, x2, x[sa[k]], x[sb[k]]};
            const double c = (sa[k] == sb[k]) ? 1.0 / 12.0 : 1.0 / 24.0;
            w[k] = c * (m + (1.0 - m) * dirichlet_power_moment(ps.ip, xs, 5));
        }
        return;
    }
    for (int k = 0; k < 6; ++k) w[k] = 0.0;
    for (int q = 0; q < 16; ++q) {
        double lam[3], wt;
        collapsed_gauss16(q, lam, wt);
        const double xq = lam[0] * x0 + lam[1] * x1 + lam[2] * x2;
        const double r = 0.5 * wt * (m + (1.0 - m) * pow(xq, ps.p));
        for (int k = 0; k < 6; ++k) w[k] += r * lam[sa[k]] * lam[sb[k]];
    }
}

// tri_sensitivity for a general exponent:  r' = p (1 - m) xi^(p-1)
template <typename T>
TM_HD void tri_sensitivity_general(const T e[3][3], const T xi[3], T m, const Material<T>& mat,
                                   const PenaltySpec& ps, T g[3]) {
    double E[3][3];
    for (int c = 0; c < 3; ++c)
        for (int d = c; d < 3; ++d) {
            E[c][d] = (double)(mat.A11 * e[c][0] * e[d][0] + mat.A22 * e[c][1] * e[d][1] +
                               mat.A12 * (e[c][0] * e[d][1] + e[c][1] * e[d][0]) + mat.A33 * e[c][2] * e[d][2]);
            E[d][c] = E[c][d];
        }
    const double x[3] = {(double)xi[0], (double)xi[1], (double)xi[2]};
    const double scale = -ps.p * (1.0 - (double)m);
    double acc[3] = {0.0, 0.0, 0.0};
    if (ps.ip > 0) {
        const double fact[4] = {1.0, 1.0, 2.0, 6.0};
        for (int i = 0; i < 3; ++i)
            for (int c = 0; c < 3; ++c)
                for (int d = 0; d < 3; ++d) {
                    int cnt[3] = {0, 0, 0};
                    cnt[c]++; cnt[d]++; cnt[i]++;
                    const double af = fact[cnt[0]] * fact[cnt[1]] * fact[cnt[2]];
                    const double xs[6] = {x[0], x[1], x[2], x[c], x[d], x[i]};
                    acc[i] += E[c][d] * af * (1.0 / 120.0) * dirichlet_power_moment(ps.ip - 1, xs, 6);
                }
    } else {
        for (int q = 0; q < 16; ++q) {
            double lam[3], wt;
            collapsed_gauss16(q, lam, wt);
            const double xq = lam[0] * x[0] + lam[1] * x[1] + lam[2] * x[2];
            double energy = 0.0;
            for (int c = 0; c < 3; ++c)
                for (int d = 0; d < 3; ++d) energy += lam[c] * lam[d] * E[c][d];
            const double f = 0.5 * wt * pow(xq, ps.p - 1.0) * energy;
            for (int i = 0; i < 3; ++i) acc[i] += f * lam[i];
        }
    }
    for (int i = 0; i < 3; ++i) g[i] = (T)(scale * acc[i]);
}

}  // namespace tmx
