// Host build of topomax_b200/csrc/tm_element.cuh (the same source the kernels use) so that
// the element arithmetic can be checked against the oracle on a machine without a GPU.
// Test infrastructure only: built by tests/test_element_host.py with g++, never shipped.
#include "../../topomax_b200/csrc/tm_element.cuh"

extern "C" {

// dense 18x18 cell matrix (row-major), cell-local dof = 2*q + comp, from vertex densities
// xi4 = (v0, v1, v2, v3) = (bottom-left, bottom-right, top-left, top-right)
void hc_cell_matrix(const double* xi4, double m, double lambda, double mu, double hx, double hy,
                    double* K) {
    auto mat = tmx::make_material<double>(lambda, mu, hx, hy);
    double wA[6], wB[6];
    tmx::moments_from_xi<double>(xi4[0], xi4[1], xi4[3], m, wA);
    tmx::moments_from_xi<double>(xi4[0], xi4[2], xi4[3], m, wB);
    for (int j = 0; j < 18; ++j) {
        double X[9][2] = {}, acc[9][2] = {};
        X[j / 2][j % 2] = 1.0;
        tmx::cell_apply<double>(X, wA, wB, mat, acc);
        for (int i = 0; i < 18; ++i) K[i * 18 + j] = acc[i / 2][i % 2];
    }
}

void hc_moments(const double* xi3, double m, double* w) {
    tmx::moments_from_xi<double>(xi3[0], xi3[1], xi3[2], m, w);
}

// cell apply with explicit moments (coarse-level form)
void hc_cell_apply_w(const double* X18, const double* wA, const double* wB, double lambda,
                     double mu, double hx, double hy, double* Y18) {
    auto mat = tmx::make_material<double>(lambda, mu, hx, hy);
    double X[9][2], acc[9][2] = {};
    for (int q = 0; q < 9; ++q) { X[q][0] = X18[2 * q]; X[q][1] = X18[2 * q + 1]; }
    tmx::cell_apply<double>(X, wA, wB, mat, acc);
    for (int q = 0; q < 9; ++q) { Y18[2 * q] = acc[q][0]; Y18[2 * q + 1] = acc[q][1]; }
}

// sensitivity contributions of one cell to its 4 vertices (v0, v1, v2, v3)
void hc_cell_sensitivity(const double* X18, const double* xi4, double m, double lambda, double mu,
                         double hx, double hy, double* g4) {
    auto mat = tmx::make_material<double>(lambda, mu, hx, hy);
    double X[9][2];
    for (int q = 0; q < 9; ++q) { X[q][0] = X18[2 * q]; X[q][1] = X18[2 * q + 1]; }
    for (int i = 0; i < 4; ++i) g4[i] = 0;
    {
        const double U[3][2] = {{X[0][0], X[0][1]}, {X[2][0], X[2][1]}, {X[8][0], X[8][1]}};
        const double M[3][2] = {{X[1][0], X[1][1]}, {X[5][0], X[5][1]}, {X[4][0], X[4][1]}};
        double e[3][3], g[3];
        const double xi[3] = {xi4[0], xi4[1], xi4[3]};
        tmx::tri_vertex_strains<double, false>(U, M, mat.kappa, e);
        tmx::tri_sensitivity<double>(e, xi, m, mat, g);
        g4[0] += g[0]; g4[1] += g[1]; g4[3] += g[2];
    }
    {
        const double U[3][2] = {{X[0][0], X[0][1]}, {X[6][0], X[6][1]}, {X[8][0], X[8][1]}};
        const double M[3][2] = {{X[3][0], X[3][1]}, {X[7][0], X[7][1]}, {X[4][0], X[4][1]}};
        double e[3][3], g[3];
        const double xi[3] = {xi4[0], xi4[2], xi4[3]};
        tmx::tri_vertex_strains<double, true>(U, M, mat.kappa, e);
        tmx::tri_sensitivity<double>(e, xi, m, mat, g);
        g4[0] += g[0]; g4[2] += g[1]; g4[3] += g[2];
    }
}

static tmx::PenaltySpec hc_spec(double p) {
    tmx::PenaltySpec ps;
    ps.p = p;
    const int ip = (int)p;
    ps.ip = ((double)ip == p && ip >= 1 && ip <= tmx::kMaxIntPenalty) ? ip : 0;
    return ps;
}

// general-exponent counterparts
void hc_moments_general(const double* xi3, double m, double p, double* w) {
    tmx::moments_general(xi3[0], xi3[1], xi3[2], m, hc_spec(p), w);
}

void hc_cell_matrix_general(const double* xi4, double m, double p, double lambda, double mu, double hx,
                            double hy, double* K) {
    auto mat = tmx::make_material<double>(lambda, mu, hx, hy);
    double wA[6], wB[6];
    tmx::moments_general(xi4[0], xi4[1], xi4[3], m, hc_spec(p), wA);
    tmx::moments_general(xi4[0], xi4[2], xi4[3], m, hc_spec(p), wB);
    for (int j = 0; j < 18; ++j) {
        double X[9][2] = {}, acc[9][2] = {};
        X[j / 2][j % 2] = 1.0;
        tmx::cell_apply<double>(X, wA, wB, mat, acc);
        for (int i = 0; i < 18; ++i) K[i * 18 + j] = acc[i / 2][i % 2];
    }
}

void hc_cell_sensitivity_general(const double* X18, const double* xi4, double m, double p, double lambda,
                                 double mu, double hx, double hy, double* g4) {
    auto mat = tmx::make_material<double>(lambda, mu, hx, hy);
    const tmx::PenaltySpec ps = hc_spec(p);
    double X[9][2];
    for (int q = 0; q < 9; ++q) { X[q][0] = X18[2 * q]; X[q][1] = X18[2 * q + 1]; }
    for (int i = 0; i < 4; ++i) g4[i] = 0;
    {
        const double U[3][2] = {{X[0][0], X[0][1]}, {X[2][0], X[2][1]}, {X[8][0], X[8][1]}};
        const double M[3][2] = {{X[1][0], X[1][1]}, {X[5][0], X[5][1]}, {X[4][0], X[4][1]}};
        double e[3][3], g[3];
        const double xi[3] = {xi4[0], xi4[1], xi4[3]};
        tmx::tri_vertex_strains<double, false>(U, M, mat.kappa, e);
        tmx::tri_sensitivity_general<double>(e, xi, m, mat, ps, g);
        g4[0] += g[0]; g4[1] += g[1]; g4[3] += g[2];
    }
    {
        const double U[3][2] = {{X[0][0], X[0][1]}, {X[6][0], X[6][1]}, {X[8][0], X[8][1]}};
        const double M[3][2] = {{X[3][0], X[3][1]}, {X[7][0], X[7][1]}, {X[4][0], X[4][1]}};
        double e[3][3], g[3];
        const double xi[3] = {xi4[0], xi4[2], xi4[3]};
        tmx::tri_vertex_strains<double, true>(U, M, mat.kappa, e);
        tmx::tri_sensitivity_general<double>(e, xi, m, mat, ps, g);
        g4[0] += g[0]; g4[2] += g[1]; g4[3] += g[2];
    }
}
}
