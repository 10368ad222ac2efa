/* CPU kernel baseline for the state solve: matrix-free penalised P2 elasticity operator and
 * Jacobi-preconditioned conjugate gradients, plain C + OpenMP, fp64.
 *
 * TEST / MEASUREMENT INFRASTRUCTURE ONLY (same rule as oracle/fem_oracle.py): nothing under
 * topomax_b200/ links or loads this; only tests/ and bench.py's CPU-baseline legs do.
 *
 * What it restates (reference = Emilinya/topomax, file:line):
 *   a(u,v) = int r(xi_h) [lambda div u div v + 2 mu eps(u):eps(v)] dx
 *                                   FEM_src/elasisity_problem.py:112-118
 *   r(xi) = m + (1-m) xi^p          src/penalizers.py:36-40
 *   homogeneous Dirichlet rows on whole fixed sides (identity rows, zero right-hand side)
 *                                   FEM_src/elasisity_problem.py:183-192, FEM_src/pde_solver.py:125
 *   mesh: RectangleMesh "right" diagonal, vector-P2 on the half-step lattice
 *                                   FEM_src/solver.py:38-49, FEM_src/elasisity_problem.py:194-196
 * The reference solves with MUMPS LU (FEM_src/pde_solver.py:130-131); a direct factorisation is
 * not feasible at the bench sizes on a host, so the like-for-like CPU *kernel* figure SURVEY.md
 * section 8(d) asks for is this Jacobi-PCG: same operator, same vectors, all host cores.
 *
 * Element arithmetic is deliberately the textbook one (numerical quadrature of B^T D B at the
 * points of a collapsed Gauss rule handed in by the caller, the rule oracle/fem_oracle.py uses),
 * NOT the closed-form moment algebra of the CUDA kernels: the two share no code and no trick.
 *
 * Layout (shared with the CUDA library and the numpy oracle):
 *   P1 fields  v = iy*(nx+1) + ix;   P2 node n = j*(2nx+1) + i, dof 2n + component.
 * Parallelisation: cell rows of equal parity touch disjoint lattice rows, so the scatter runs
 * as two race-free sweeps (even rows, odd rows) with one thread per cell row.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define OC_MAX_NQ 64

typedef struct {
    int nx, ny;          /* cells */
    double W, H;         /* rectangle size */
    double lam, mu;      /* Lame parameters */
    double p, m;         /* SIMP exponent and minimum */
    int fixed[4];        /* Left, Right, Bottom, Top */
    int nq;              /* quadrature points per triangle */
    const double* pts;   /* nq x 3 barycentric points */
    const double* wts;   /* nq weights summing to 1 */
} oc_problem;

/* strain tables of the two triangle types (all triangles of a type are translates):
 * B[t][q] maps the 12 local dofs (node-major) to (e_xx, e_yy, gamma_xy) */
typedef struct {
    double B[2][OC_MAX_NQ][3][12];
    double area[2];
    int dn[2][6][2]; /* lattice offsets (di, dj) of the six local nodes from the cell origin */
} oc_tables;

static const int kEdges[3][2] = {{0, 1}, {1, 2}, {0, 2}};

static void build_tables(const oc_problem* P, oc_tables* T) {
    const double hx = P->W / P->nx, hy = P->H / P->ny;
    /* vertices of T_A = (v0, v1, v3), T_B = (v0, v2, v3) in lattice steps */
    const int vs[2][3][2] = {{{0, 0}, {2, 0}, {2, 2}}, {{0, 0}, {0, 2}, {2, 2}}};
    for (int t = 0; t < 2; ++t) {
        double X[3][2];
        for (int a = 0; a < 3; ++a) {
            X[a][0] = 0.5 * hx * vs[t][a][0];
            X[a][1] = 0.5 * hy * vs[t][a][1];
            T->dn[t][a][0] = vs[t][a][0];
            T->dn[t][a][1] = vs[t][a][1];
        }
        for (int k = 0; k < 3; ++k) {
            T->dn[t][3 + k][0] = (vs[t][kEdges[k][0]][0] + vs[t][kEdges[k][1]][0]) / 2;
            T->dn[t][3 + k][1] = (vs[t][kEdges[k][0]][1] + vs[t][kEdges[k][1]][1]) / 2;
        }
        /* J = [x1-x0, x2-x0]; grad lam_1, lam_2 = rows of J^-1; lam_0 = 1 - lam_1 - lam_2 */
        const double a = X[1][0] - X[0][0], b = X[2][0] - X[0][0];
        const double c = X[1][1] - X[0][1], d = X[2][1] - X[0][1];
        const double det = a * d - b * c;
        T->area[t] = 0.5 * fabs(det);
        double gl[3][2];
        gl[1][0] = d / det;  gl[1][1] = -b / det;
        gl[2][0] = -c / det; gl[2][1] = a / det;
        gl[0][0] = -(gl[1][0] + gl[2][0]);
        gl[0][1] = -(gl[1][1] + gl[2][1]);
        for (int q = 0; q < P->nq; ++q) {
            const double* l = P->pts + 3 * q;
            double g[6][2];
            for (int v = 0; v < 3; ++v)
                for (int k = 0; k < 2; ++k) g[v][k] = (4.0 * l[v] - 1.0) * gl[v][k];
            for (int e = 0; e < 3; ++e) {
                const int v0 = kEdges[e][0], v1 = kEdges[e][1];
                for (int k = 0; k < 2; ++k) g[3 + e][k] = 4.0 * (l[v0] * gl[v1][k] + l[v1] * gl[v0][k]);
            }
            memset(T->B[t][q], 0, sizeof(T->B[t][q]));
            for (int n = 0; n < 6; ++n) {
                T->B[t][q][0][2 * n] = g[n][0];
                T->B[t][q][1][2 * n + 1] = g[n][1];
                T->B[t][q][2][2 * n] = g[n][1];
                T->B[t][q][2][2 * n + 1] = g[n][0];
            }
        }
    }
}

static inline int node_fixed(const oc_problem* P, int i, int j) {
    return (P->fixed[0] && i == 0) || (P->fixed[1] && i == 2 * P->nx) ||
           (P->fixed[2] && j == 0) || (P->fixed[3] && j == 2 * P->ny);
}

static inline double simp(const oc_problem* P, double x) {
    const double xp = (P->p == 3.0) ? x * x * x : pow(x, P->p);
    return P->m + (1.0 - P->m) * xp;
}

/* one cell row: y += K_row x (mode 0) or diag += diag(K_row) (mode 1) */
__attribute__((target_clones("default", "avx2")))
static void cell_row(const oc_problem* P, const oc_tables* T, int iy, const double* xi, const double* x,
                     double* y, int mode) {
    const int nx = P->nx, Lx = 2 * nx + 1, n1x = nx + 1;
    const double lam = P->lam, mu = P->mu;
    for (int ix = 0; ix < nx; ++ix) {
        const int v0 = iy * n1x + ix;
        const double xv[2][3] = {{xi[v0], xi[v0 + 1], xi[v0 + n1x + 1]}, {xi[v0], xi[v0 + n1x], xi[v0 + n1x + 1]}};
        for (int t = 0; t < 2; ++t) {
            int dof[12], fixedn[6];
            double xe[12], ye[12];
            for (int n = 0; n < 6; ++n) {
                const int i = 2 * ix + T->dn[t][n][0], j = 2 * iy + T->dn[t][n][1];
                const int node = j * Lx + i;
                fixedn[n] = node_fixed(P, i, j);
                dof[2 * n] = 2 * node;
                dof[2 * n + 1] = 2 * node + 1;
                if (mode == 0) {
                    xe[2 * n] = fixedn[n] ? 0.0 : x[2 * node];
                    xe[2 * n + 1] = fixedn[n] ? 0.0 : x[2 * node + 1];
                }
                ye[2 * n] = ye[2 * n + 1] = 0.0;
            }
            for (int q = 0; q < P->nq; ++q) {
                const double* l = P->pts + 3 * q;
                const double xq = l[0] * xv[t][0] + l[1] * xv[t][1] + l[2] * xv[t][2];
                const double cq = simp(P, xq) * P->wts[q] * T->area[t];
                const double(*B)[12] = T->B[t][q];
                if (mode == 0) {
                    double e0 = 0.0, e1 = 0.0, e2 = 0.0;
                    for (int k = 0; k < 12; ++k) {
                        e0 += B[0][k] * xe[k];
                        e1 += B[1][k] * xe[k];
                        e2 += B[2][k] * xe[k];
                    }
                    const double s0 = cq * ((lam + 2.0 * mu) * e0 + lam * e1);
                    const double s1 = cq * (lam * e0 + (lam + 2.0 * mu) * e1);
                    const double s2 = cq * mu * e2;
                    for (int k = 0; k < 12; ++k) ye[k] += B[0][k] * s0 + B[1][k] * s1 + B[2][k] * s2;
                } else {
                    for (int k = 0; k < 12; ++k) {
                        const double b0 = B[0][k], b1 = B[1][k], b2 = B[2][k];
                        ye[k] += cq * ((lam + 2.0 * mu) * (b0 * b0 + b1 * b1) + 2.0 * lam * b0 * b1 + mu * b2 * b2);
                    }
                }
            }
            for (int n = 0; n < 6; ++n) {
                if (fixedn[n]) continue;
                y[dof[2 * n]] += ye[2 * n];
                y[dof[2 * n + 1]] += ye[2 * n + 1];
            }
        }
    }
}

static void sweep(const oc_problem* P, const oc_tables* T, const double* xi, const double* x, double* y,
                  int mode) {
    const int Lx = 2 * P->nx + 1, Ly = 2 * P->ny + 1;
    const long nu = 2L * Lx * Ly;
#pragma omp parallel for schedule(static)
    for (long k = 0; k < nu; ++k) y[k] = 0.0;
    for (int parity = 0; parity < 2; ++parity) {
#pragma omp parallel for schedule(dynamic, 4)
        for (int iy = parity; iy < P->ny; iy += 2) cell_row(P, T, iy, xi, x, y, mode);
    }
    /* identity rows on the fixed sides */
#pragma omp parallel for schedule(static)
    for (int j = 0; j < Ly; ++j)
        for (int i = 0; i < Lx; ++i)
            if (node_fixed(P, i, j)) {
                const long n = (long)j * Lx + i;
                y[2 * n] = mode == 0 ? x[2 * n] : 1.0;
                y[2 * n + 1] = mode == 0 ? x[2 * n + 1] : 1.0;
            }
}

int oc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void oc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* y = A x with A = P K(xi) P + (I - P), P the projector on the free dofs */
int oc_elast_apply(const oc_problem* P, const double* xi, const double* x, double* y) {
    if (P->nq < 1 || P->nq > OC_MAX_NQ) return 1;
    oc_tables* T = (oc_tables*)malloc(sizeof(oc_tables));
    if (!T) return 2;
    build_tables(P, T);
    sweep(P, T, xi, x, y, 0);
    free(T);
    return 0;
}

int oc_elast_diag(const oc_problem* P, const double* xi, double* diag) {
    if (P->nq < 1 || P->nq > OC_MAX_NQ) return 1;
    oc_tables* T = (oc_tables*)malloc(sizeof(oc_tables));
    if (!T) return 2;
    build_tables(P, T);
    sweep(P, T, xi, NULL, diag, 1);
    free(T);
    return 0;
}

static double now_seconds(void) {
#ifdef _OPENMP
    return omp_get_wtime();
#else
    return 0.0;
#endif
}

/* Jacobi-PCG for A u = b (b's fixed dofs are treated as 0), zero initial guess.
 * Stops when ||r||_2 <= rtol ||b||_2 or after maxit iterations.
 * seconds = wall time of the iteration loop only (set-up of the diagonal excluded). */
int oc_jacobi_pcg(const oc_problem* P, const double* xi, const double* b, double* u, double rtol, int maxit,
                  int* iters, double* relres, double* seconds) {
    if (P->nq < 1 || P->nq > OC_MAX_NQ) return 1;
    const int Lx = 2 * P->nx + 1, Ly = 2 * P->ny + 1;
    const long nu = 2L * Lx * Ly;
    oc_tables* T = (oc_tables*)malloc(sizeof(oc_tables));
    double* dinv = (double*)malloc(sizeof(double) * nu);
    double* r = (double*)malloc(sizeof(double) * nu);
    double* p = (double*)malloc(sizeof(double) * nu);
    double* Ap = (double*)malloc(sizeof(double) * nu);
    if (!T || !dinv || !r || !p || !Ap) {
        free(T); free(dinv); free(r); free(p); free(Ap);
        return 2;
    }
    build_tables(P, T);
    sweep(P, T, xi, NULL, dinv, 1);
    double bb = 0.0, rz = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : bb, rz)
    for (long n = 0; n < (long)Lx * Ly; ++n) {
        const int fx = node_fixed(P, (int)(n % Lx), (int)(n / Lx));
        for (int c = 0; c < 2; ++c) {
            const long k = 2 * n + c;
            dinv[k] = 1.0 / dinv[k];
            r[k] = fx ? 0.0 : b[k];
            u[k] = 0.0;
            p[k] = dinv[k] * r[k];
            bb += r[k] * r[k];
            rz += r[k] * p[k];
        }
    }
    int it = 0;
    double rr = bb;
    const double t0 = now_seconds();
    while (it < maxit && bb > 0.0 && sqrt(rr) > rtol * sqrt(bb)) {
        sweep(P, T, xi, p, Ap, 0);
        double pAp = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : pAp)
        for (long k = 0; k < nu; ++k) pAp += p[k] * Ap[k];
        const double alpha = rz / pAp;
        double rz_new = 0.0;
        rr = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : rz_new, rr)
        for (long k = 0; k < nu; ++k) {
            u[k] += alpha * p[k];
            const double rk = r[k] - alpha * Ap[k];
            r[k] = rk;
            rr += rk * rk;
            rz_new += rk * dinv[k] * rk;
        }
        const double beta = rz_new / rz;
        rz = rz_new;
#pragma omp parallel for schedule(static)
        for (long k = 0; k < nu; ++k) p[k] = dinv[k] * r[k] + beta * p[k];
        ++it;
    }
    *seconds = now_seconds() - t0;
    *iters = it;
    *relres = bb > 0.0 ? sqrt(rr / bb) : 0.0;
    free(T); free(dinv); free(r); free(p); free(Ap);
    return 0;
}
