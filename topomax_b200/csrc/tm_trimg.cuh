// Geometric multigrid on per-triangle local matrices, for the fluid row's preconditioner
// (one V-cycle per velocity component on M_r + K, and a V-cycle on the P1
// "Darcy" Laplacian int (1/r) grad.grad for the pressure block).
//
// On the nested right-diagonal meshes (cells halved per level) both the P2 and the P1 spaces are
// nested, and a coarse basis function restricted to a coarse triangle is a polynomial of the same
// degree on each of its four children.  The Galerkin coarse operator P^T A P is therefore again a sum
// of per-triangle local matrices,  L_coarse(T) = sum_children C^T L_child C,  with C the constant
// NODES x NODES interpolation matrix of the child position.  Every level is thus the same object --
// NODES(NODES+1)/2 numbers per triangle -- and one set of work items serves both hierarchies:
// NODES = 6 (P2 lattice, `ncomp` interleaved components, optional Dirichlet boundary) and
// NODES = 3 (P1 vertices).  Prolongation is a per-fine-node gather of the coarse polynomial,
// restriction its exact transpose (scatter), smoothing Chebyshev-Jacobi, the coarsest level an
// explicit inverse built on the host.
//
// As in tm_fluid.cuh the work items and the V-cycle are host-compilable and are checked on the CPU
// against scipy Galerkin products (tests/hostcheck/fluid_host.cpp); the CUDA back-end adds launches.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <utility>
#include <vector>

#include "tm_element.cuh"  // TM_HD

namespace tmx {

struct TriLevel {
    int nx, ny;          // cells
    int ncomp;           // interleaved components sharing the scalar operator
    int fixed_boundary;  // NODES = 6 only: all boundary lattice nodes are Dirichlet (identity rows)
};

template <int NODES>
TM_HD size_t trimg_num_nodes(const TriLevel& g) {
    return NODES == 6 ? (size_t)(2 * g.nx + 1) * (2 * g.ny + 1) : (size_t)(g.nx + 1) * (g.ny + 1);
}
TM_HD size_t trimg_num_tri(const TriLevel& g) { return (size_t)2 * g.nx * g.ny; }

template <int NODES>
TM_HD int trimg_sym(int i, int j) {
    const int a = i < j ? i : j, b = i < j ? j : i;
    return a * NODES - a * (a - 1) / 2 + (b - a);
}

// nodes of triangle (cx, cy, t): T_A = (v0, v1, v3), T_B = (v0, v2, v3); P2 local order
// 0,1,2 vertices, 3 = mid(0,1), 4 = mid(1,2), 5 = mid(0,2)
template <int NODES>
TM_HD void trimg_nodes(const TriLevel& g, int cx, int cy, int t, int (&node)[NODES], bool (&interior)[NODES]) {
    if (NODES == 6) {
        const int Lx = 2 * g.nx + 1, Ly = 2 * g.ny + 1;
#pragma unroll
        for (int k = 0; k < NODES; ++k) {
            // lattice offsets: A {0,2,2,1,2,1 | 0,0,2,0,1,1}, B {0,0,2,0,1,1 | 0,2,2,1,2,1}
            const int da = k == 0 ? 0 : (k == 1 || k == 2 || k == 4) ? 2 : 1;                 // A: di ; B: dj
            const int db = (k == 0 || k == 1 || k == 3) ? 0 : (k == 2) ? 2 : 1;                // A: dj ; B: di
            const int i = 2 * cx + (t == 0 ? da : db), j = 2 * cy + (t == 0 ? db : da);
            node[k] = j * Lx + i;
            interior[k] = !g.fixed_boundary || (i > 0 && j > 0 && i < Lx - 1 && j < Ly - 1);
        }
    } else {
        const int v0 = cy * (g.nx + 1) + cx;
        node[0] = v0;
        node[1] = t == 0 ? v0 + 1 : v0 + g.nx + 1;
        node[2] = v0 + g.nx + 2;
#pragma unroll
        for (int k = 0; k < NODES; ++k) interior[k] = true;
    }
}

TM_HD void trimg_tid_to_cell(const TriLevel& g, size_t tid, int& cx, int& cy, int& t) {
    t = (int)(tid & 1);
    const size_t cell = tid >> 1;
    cy = (int)(cell / g.nx);
    cx = (int)(cell - (size_t)cy * g.nx);
}

// basis values at barycentric point l
template <int NODES>
TM_HD void trimg_basis(const double (&l)[3], double (&phi)[NODES]) {
    if (NODES == 6) {
        phi[0] = l[0] * (2 * l[0] - 1);
        phi[1] = l[1] * (2 * l[1] - 1);
        phi[2] = l[2] * (2 * l[2] - 1);
        phi[3 % NODES] = 4 * l[0] * l[1];
        phi[4 % NODES] = 4 * l[1] * l[2];
        phi[5 % NODES] = 4 * l[0] * l[2];
    } else {
        phi[0] = l[0];
        phi[1] = l[1];
        phi[2] = l[2];
    }
}

// y += A x on one triangle (Dirichlet columns dropped, Dirichlet rows not written)
template <int NODES, class Add>
TM_HD void trimg_body_apply(const TriLevel& g, const double* Lm, size_t ntri, const double* x, double* y,
                            size_t tid, Add add) {
    int cx, cy, t, node[NODES];
    bool in[NODES];
    trimg_tid_to_cell(g, tid, cx, cy, t);
    trimg_nodes<NODES>(g, cx, cy, t, node, in);
    double L[NODES * (NODES + 1) / 2];
#pragma unroll
    for (int k = 0; k < NODES * (NODES + 1) / 2; ++k) L[k] = Lm[(size_t)k * ntri + tid];
    for (int c = 0; c < g.ncomp; ++c) {
        double xv[NODES];
#pragma unroll
        for (int k = 0; k < NODES; ++k) xv[k] = in[k] ? x[(size_t)node[k] * g.ncomp + c] : 0.0;
#pragma unroll
        for (int i = 0; i < NODES; ++i) {
            if (!in[i]) continue;
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < NODES; ++j) s += L[trimg_sym<NODES>(i, j)] * xv[j];
            add(&y[(size_t)node[i] * g.ncomp + c], s);
        }
    }
}

template <int NODES, class Add>
TM_HD void trimg_body_diag(const TriLevel& g, const double* Lm, size_t ntri, double* diag, size_t tid, Add add) {
    int cx, cy, t, node[NODES];
    bool in[NODES];
    trimg_tid_to_cell(g, tid, cx, cy, t);
    trimg_nodes<NODES>(g, cx, cy, t, node, in);
#pragma unroll
    for (int i = 0; i < NODES; ++i)
        if (in[i]) {
            const double a = Lm[(size_t)trimg_sym<NODES>(i, i) * ntri + tid];
            for (int c = 0; c < g.ncomp; ++c) add(&diag[(size_t)node[i] * g.ncomp + c], a);
        }
}

// children of a coarse triangle: fine-cell offsets inside the 2x2 block and the child's type
//   T_A: (0,0,A) (1,0,A) (1,0,B) (1,1,A)      T_B: (0,0,B) (0,1,A) (0,1,B) (1,1,B)
TM_HD void trimg_child(int tc, int child, int& dx, int& dy, int& tf) {
    const int cdx[2][4] = {{0, 1, 1, 1}, {0, 0, 0, 1}};
    const int cdy[2][4] = {{0, 0, 0, 1}, {0, 1, 1, 1}};
    const int ctf[2][4] = {{0, 0, 1, 0}, {1, 0, 1, 1}};
    dx = cdx[tc][child];
    dy = cdy[tc][child];
    tf = ctf[tc][child];
}

// interpolation matrices C[coarse type][child][child node k][coarse node K] = phi_K(child node k)
template <int NODES>
struct TriCoarsenTable {
    double C[2][4][NODES][NODES];
};

template <int NODES>
inline TriCoarsenTable<NODES> trimg_make_coarsen_table() {
    TriCoarsenTable<NODES> tab;
    // coordinates in fine-cell units inside the 2x2 block; triangle vertices per type
    const double vtx[2][3][2] = {{{0, 0}, {1, 0}, {1, 1}}, {{0, 0}, {0, 1}, {1, 1}}};  // of a unit cell
    for (int tc = 0; tc < 2; ++tc)
        for (int ch = 0; ch < 4; ++ch) {
            int dx, dy, tf;
            trimg_child(tc, ch, dx, dy, tf);
            // child's nodes: vertices then (for P2) the three mids
            double pts[6][2];
            for (int v = 0; v < 3; ++v) {
                pts[v][0] = dx + vtx[tf][v][0];
                pts[v][1] = dy + vtx[tf][v][1];
            }
            const int ea[3] = {0, 1, 0}, eb[3] = {1, 2, 2};
            for (int e = 0; e < 3; ++e) {
                pts[3 + e][0] = 0.5 * (pts[ea[e]][0] + pts[eb[e]][0]);
                pts[3 + e][1] = 0.5 * (pts[ea[e]][1] + pts[eb[e]][1]);
            }
            for (int k = 0; k < NODES; ++k) {
                // barycentrics in the coarse triangle (vertices 2*vtx[tc])
                const double a = pts[k][0] / 2.0, b = pts[k][1] / 2.0;  // coarse-cell coordinates in [0,1]
                double l[3];
                if (tc == 0) { l[0] = 1 - a; l[1] = a - b; l[2] = b; }
                else { l[0] = 1 - b; l[1] = b - a; l[2] = a; }
                double phi[NODES];
                trimg_basis<NODES>(l, phi);
                for (int K = 0; K < NODES; ++K) tab.C[tc][ch][k][K] = phi[K];
            }
        }
    return tab;
}

// coarse local matrix of triangle tid_c = sum over its 4 children of C^T L_child C
template <int NODES>
TM_HD void trimg_body_coarsen(const TriCoarsenTable<NODES>& tab, const TriLevel& gf, const double* Lf, size_t ntri_f,
                              const TriLevel& gc, double* Lc, size_t ntri_c, size_t tid_c) {
    constexpr int NE = NODES * (NODES + 1) / 2;
    int cx, cy, tc;
    trimg_tid_to_cell(gc, tid_c, cx, cy, tc);
    double acc[NE];
#pragma unroll
    for (int k = 0; k < NE; ++k) acc[k] = 0.0;
    for (int ch = 0; ch < 4; ++ch) {
        int dx, dy, tf;
        trimg_child(tc, ch, dx, dy, tf);
        const size_t tid_f = 2 * ((size_t)(2 * cy + dy) * gf.nx + (2 * cx + dx)) + tf;
        double L[NE];
#pragma unroll
        for (int k = 0; k < NE; ++k) L[k] = Lf[(size_t)k * ntri_f + tid_f];
        // T = L C  (NODES x NODES), then acc += C^T T (upper triangle)
        double T[NODES][NODES];
        for (int i = 0; i < NODES; ++i)
            for (int K = 0; K < NODES; ++K) {
                double s = 0.0;
                for (int j = 0; j < NODES; ++j) s += L[trimg_sym<NODES>(i, j)] * tab.C[tc][ch][j][K];
                T[i][K] = s;
            }
        int k = 0;
        for (int I = 0; I < NODES; ++I)
            for (int K = I; K < NODES; ++K, ++k) {
                double s = 0.0;
                for (int i = 0; i < NODES; ++i) s += tab.C[tc][ch][i][I] * T[i][K];
                acc[k] += s;
            }
    }
#pragma unroll
    for (int k = 0; k < NE; ++k) Lc[(size_t)k * ntri_c + tid_c] = acc[k];
}

// coarse nodes and weights of the coarse polynomial at fine node `nf` (index into the fine node
// array); returns false for a fine Dirichlet node.  cin[k]: coarse node is not Dirichlet.
template <int NODES>
TM_HD bool trimg_transfer_stencil(const TriLevel& gf, const TriLevel& gc, size_t nf, int (&cnode)[NODES],
                                  double (&w)[NODES], bool (&cin)[NODES]) {
    const int span = NODES == 6 ? 4 : 2;  // fine node steps per coarse cell
    const int Lxf = NODES == 6 ? 2 * gf.nx + 1 : gf.nx + 1, Lyf = NODES == 6 ? 2 * gf.ny + 1 : gf.ny + 1;
    const int j = (int)(nf / Lxf), i = (int)(nf - (size_t)j * Lxf);
    if (NODES == 6 && gf.fixed_boundary && (i == 0 || j == 0 || i == Lxf - 1 || j == Lyf - 1)) return false;
    int I = i / span, J = j / span;
    if (I > gc.nx - 1) I = gc.nx - 1;
    if (J > gc.ny - 1) J = gc.ny - 1;
    const double a = (double)(i - span * I) / span, b = (double)(j - span * J) / span;
    const int t = a >= b ? 0 : 1;
    double l[3];
    if (t == 0) { l[0] = 1 - a; l[1] = a - b; l[2] = b; }
    else { l[0] = 1 - b; l[1] = b - a; l[2] = a; }
    trimg_basis<NODES>(l, w);
    trimg_nodes<NODES>(gc, I, J, t, cnode, cin);
    return true;
}

// x_f += P x_c   (gather, one work item per fine node)
template <int NODES>
TM_HD void trimg_body_prolong_add(const TriLevel& gf, const TriLevel& gc, const double* xc, double* xf, size_t nf) {
    int cnode[NODES];
    double w[NODES];
    bool cin[NODES];
    if (!trimg_transfer_stencil<NODES>(gf, gc, nf, cnode, w, cin)) return;
    for (int c = 0; c < gf.ncomp; ++c) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < NODES; ++k)
            if (cin[k] && w[k] != 0.0) s += w[k] * xc[(size_t)cnode[k] * gf.ncomp + c];
        xf[nf * gf.ncomp + c] += s;
    }
}

// r_c += P^T r_f   (scatter, one work item per fine node)
template <int NODES, class Add>
TM_HD void trimg_body_restrict(const TriLevel& gf, const TriLevel& gc, const double* rf, double* rc, size_t nf,
                               Add add) {
    int cnode[NODES];
    double w[NODES];
    bool cin[NODES];
    if (!trimg_transfer_stencil<NODES>(gf, gc, nf, cnode, w, cin)) return;
    for (int c = 0; c < gf.ncomp; ++c) {
        const double v = rf[nf * gf.ncomp + c];
#pragma unroll
        for (int k = 0; k < NODES; ++k)
            if (cin[k] && w[k] != 0.0) add(&rc[(size_t)cnode[k] * gf.ncomp + c], w[k] * v);
    }
}


// ---------------------------------------------------------------------------------------------
// Node -> incident triangles (for gather-type kernels: every node written once, fixed summation
// order).  Returns the number of incident (triangle id, local node index) pairs of node `n`.
//   vertex (ix, iy):   cell (ix, iy) A:0 B:0 | cell (ix-1, iy) A:1 | cell (ix, iy-1) B:1 |
//                      cell (ix-1, iy-1) A:2 B:2
//   P2 lattice node (i, j), by parity:  (even, even) the vertex rule;  (odd, even) horizontal edge:
//   cell below B:4, cell above A:3;  (even, odd) vertical edge: cell left A:4, cell right B:3;
//   (odd, odd) the diagonal of its cell: A:5, B:5.
// ---------------------------------------------------------------------------------------------
template <int NODES>
TM_HD int trimg_incident(const TriLevel& g, size_t n, size_t (&tri)[6], int (&local)[6]) {
    int cnt = 0;
    auto push = [&](int cx, int cy, int t, int k) {
        if (cx < 0 || cy < 0 || cx >= g.nx || cy >= g.ny) return;
        tri[cnt] = 2 * ((size_t)cy * g.nx + cx) + t;
        local[cnt] = k;
        ++cnt;
    };
    int ix, iy;
    bool vertex = true;
    if (NODES == 6) {
        const int Lx = 2 * g.nx + 1;
        const int j = (int)(n / Lx), i = (int)(n - (size_t)j * Lx);
        const bool io = i & 1, jo = j & 1;
        ix = i >> 1;
        iy = j >> 1;
        if (io && !jo) {  // horizontal edge between vertices ix and ix+1 on vertex row iy
            vertex = false;
            push(ix, iy - 1, 1, 4);
            push(ix, iy, 0, 3);
        } else if (!io && jo) {  // vertical edge between vertex rows iy and iy+1 on vertex column ix
            vertex = false;
            push(ix - 1, iy, 0, 4);
            push(ix, iy, 1, 3);
        } else if (io && jo) {
            vertex = false;
            push(ix, iy, 0, 5);
            push(ix, iy, 1, 5);
        }
    } else {
        iy = (int)(n / (g.nx + 1));
        ix = (int)(n - (size_t)iy * (g.nx + 1));
    }
    if (vertex) {
        push(ix, iy, 0, 0);
        push(ix, iy, 1, 0);
        push(ix - 1, iy, 0, 1);
        push(ix, iy - 1, 1, 1);
        push(ix - 1, iy - 1, 0, 2);
        push(ix - 1, iy - 1, 1, 2);
    }
    return cnt;
}

// y[n] = (A x)[n] by gathering over the incident triangles (deterministic; Dirichlet rows give 0)
template <int NODES>
TM_HD void trimg_body_apply_gather(const TriLevel& g, const double* Lm, size_t ntri, const double* x, double* y,
                                   size_t n) {
    size_t tri[6];
    int local[6];
    const int cnt = trimg_incident<NODES>(g, n, tri, local);
    double acc[2] = {0.0, 0.0};  // ncomp <= 2
    bool fixed_node = false;
    for (int q = 0; q < cnt; ++q) {
        int cx, cy, t, node[NODES];
        bool in[NODES];
        trimg_tid_to_cell(g, tri[q], cx, cy, t);
        trimg_nodes<NODES>(g, cx, cy, t, node, in);
        const int i = local[q];
        if (!in[i]) {
            fixed_node = true;
            break;
        }
        for (int c = 0; c < g.ncomp; ++c) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < NODES; ++j)
                if (in[j]) s += Lm[(size_t)trimg_sym<NODES>(i, j) * ntri + tri[q]] * x[(size_t)node[j] * g.ncomp + c];
            acc[c] += s;
        }
    }
    for (int c = 0; c < g.ncomp; ++c) y[n * g.ncomp + c] = fixed_node ? 0.0 : acc[c];
}


// diagonal by gather (Dirichlet nodes: 1)
template <int NODES>
TM_HD void trimg_body_diag_gather(const TriLevel& g, const double* Lm, size_t ntri, double* diag, size_t n) {
    size_t tri[6];
    int local[6];
    const int cnt = trimg_incident<NODES>(g, n, tri, local);
    double acc = 0.0;
    bool fixed_node = false;
    for (int q = 0; q < cnt; ++q) {
        int cx, cy, t, node[NODES];
        bool in[NODES];
        trimg_tid_to_cell(g, tri[q], cx, cy, t);
        trimg_nodes<NODES>(g, cx, cy, t, node, in);
        if (!in[local[q]]) {
            fixed_node = true;
            break;
        }
        acc += Lm[(size_t)trimg_sym<NODES>(local[q], local[q]) * ntri + tri[q]];
    }
    for (int c = 0; c < g.ncomp; ++c) diag[n * g.ncomp + c] = fixed_node ? 1.0 : acc;
}

// r_c[nc] = (P^T r_f)[nc] by gather: the fine nodes whose transfer stencil can contain coarse node nc
// lie within one coarse cell of it; each is asked for its stencil (the same function the scatter form
// and the prolongation use), so this is the exact transpose with a fixed summation order.
template <int NODES>
TM_HD void trimg_body_restrict_gather(const TriLevel& gf, const TriLevel& gc, const double* rf, double* rc,
                                      size_t nc) {
    const int span = NODES == 6 ? 4 : 2;                  // fine node steps per coarse cell
    const int cstep = NODES == 6 ? 2 : 2;                 // fine node steps per coarse NODE step
    const int Lxc = NODES == 6 ? 2 * gc.nx + 1 : gc.nx + 1;
    const int Lxf = NODES == 6 ? 2 * gf.nx + 1 : gf.nx + 1, Lyf = NODES == 6 ? 2 * gf.ny + 1 : gf.ny + 1;
    const int jc = (int)(nc / Lxc), ic = (int)(nc - (size_t)jc * Lxc);
    const int i0 = ic * cstep, j0 = jc * cstep;           // the coarse node on the fine lattice
    double acc[2] = {0.0, 0.0};
    for (int j = j0 - span; j <= j0 + span; ++j) {
        if (j < 0 || j >= Lyf) continue;
        for (int i = i0 - span; i <= i0 + span; ++i) {
            if (i < 0 || i >= Lxf) continue;
            int cnode[NODES];
            double w[NODES];
            bool cin[NODES];
            const size_t nf = (size_t)j * Lxf + i;
            if (!trimg_transfer_stencil<NODES>(gf, gc, nf, cnode, w, cin)) continue;
#pragma unroll
            for (int k = 0; k < NODES; ++k)
                if ((size_t)cnode[k] == nc && cin[k] && w[k] != 0.0)
                    for (int c = 0; c < gf.ncomp; ++c) acc[c] += w[k] * rf[nf * gf.ncomp + c];
        }
    }
    for (int c = 0; c < gf.ncomp; ++c) rc[nc * gf.ncomp + c] = acc[c];
}

// ---------------------------------------------------------------------------------------------
// Coarsest level: explicit inverse of the assembled scalar operator, built on the host from the
// level's local matrices (both back-ends).  Dirichlet rows (empty after assembly) become identity.
// ---------------------------------------------------------------------------------------------
struct TriSerialAdd {
    TM_HD void operator()(double* p, double v) const { *p += v; }
};

inline std::vector<double> trimg_dense_inverse(std::vector<double> A, int n) {  // Gauss-Jordan, partial pivoting
    std::vector<double> inv((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) inv[(size_t)i * n + i] = 1.0;
    for (int c = 0; c < n; ++c) {
        int piv = c;
        for (int r = c + 1; r < n; ++r)
            if (std::fabs(A[(size_t)r * n + c]) > std::fabs(A[(size_t)piv * n + c])) piv = r;
        if (piv != c)
            for (int k = 0; k < n; ++k) {
                std::swap(A[(size_t)c * n + k], A[(size_t)piv * n + k]);
                std::swap(inv[(size_t)c * n + k], inv[(size_t)piv * n + k]);
            }
        const double d = 1.0 / A[(size_t)c * n + c];
        for (int k = 0; k < n; ++k) {
            A[(size_t)c * n + k] *= d;
            inv[(size_t)c * n + k] *= d;
        }
        for (int r = 0; r < n; ++r) {
            if (r == c) continue;
            const double f = A[(size_t)r * n + c];
            if (f == 0.0) continue;
            for (int k = 0; k < n; ++k) {
                A[(size_t)r * n + k] -= f * A[(size_t)c * n + k];
                inv[(size_t)r * n + k] -= f * inv[(size_t)c * n + k];
            }
        }
    }
    return inv;
}

// Lm: the coarsest level's local matrices on the HOST; returns the n x n inverse (n = scalar nodes)
template <int NODES>
inline std::vector<double> trimg_coarse_inverse(const TriLevel& g, const double* Lm, int& n_out) {
    const TriLevel one{g.nx, g.ny, 1, g.fixed_boundary};
    const int n = (int)trimg_num_nodes<NODES>(one);
    const size_t nt = trimg_num_tri(one);
    std::vector<double> A((size_t)n * n, 0.0), e(n), col(n);
    for (int j = 0; j < n; ++j) {
        std::fill(e.begin(), e.end(), 0.0);
        e[j] = 1.0;
        std::fill(col.begin(), col.end(), 0.0);
        for (size_t t = 0; t < nt; ++t) trimg_body_apply<NODES>(one, Lm, nt, e.data(), col.data(), t, TriSerialAdd{});
        for (int i = 0; i < n; ++i) A[(size_t)i * n + j] = col[i];
    }
    for (int i = 0; i < n; ++i)
        if (A[(size_t)i * n + i] == 0.0) A[(size_t)i * n + i] = 1.0;
    n_out = n;
    return trimg_dense_inverse(std::move(A), n);
}

// ---------------------------------------------------------------------------------------------
// V-cycle over a back-end BK (the same code on the host and on the CUDA back-end).  BK supplies
// per level l: n(l), residual(l, b, x, r) [r = b - A x], cheb_first(l, b, d, x, s) [d = x = s D^-1 b],
// cheb_next(l, r, d, x, c1, c2) [d = c1 d + c2 D^-1 r; x += d], restrict_to(l, r_fine, b_coarse),
// prolong_add(l, x_coarse, x_fine), coarse_solve(b, x), add(l, e, x) [x += e], lmax(l) and the work
// vectors vec(l, which).
// ---------------------------------------------------------------------------------------------
enum { TRIMG_B = 0, TRIMG_X = 1, TRIMG_R = 2, TRIMG_D = 3, TRIMG_E = 4, TRIMG_NVEC = 5 };

struct TriMGParams {
    int fine_steps = 2, coarse_steps = 3;
    double ratio = 30.0;  // smoothing interval [lmax/ratio, lmax]
};

// x = q(A) b from a zero guess, `steps` Chebyshev-Jacobi steps
template <class BK>
void trimg_cheb(BK& bk, int l, typename BK::Vec b, typename BK::Vec x, typename BK::Vec d, typename BK::Vec r,
                int steps, double ratio) {
    const double lmax = bk.lmax(l), lmin = lmax / ratio;
    const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
    double rho_old = 1.0 / sigma;
    bk.cheb_first(l, b, d, x, 1.0 / theta);
    for (int k = 1; k < steps; ++k) {
        const double rho = 1.0 / (2.0 * sigma - rho_old);
        bk.residual(l, b, x, r);
        bk.cheb_next(l, r, d, x, rho * rho_old, 2.0 * rho / delta);
        rho_old = rho;
    }
}

template <class BK>
void trimg_vcycle(BK& bk, int l, const TriMGParams& prm) {
    typename BK::Vec b = bk.vec(l, TRIMG_B), x = bk.vec(l, TRIMG_X), r = bk.vec(l, TRIMG_R), d = bk.vec(l, TRIMG_D),
                     e = bk.vec(l, TRIMG_E);
    if (l == bk.levels() - 1) {
        bk.coarse_solve(b, x);
        return;
    }
    const int steps = l == 0 ? prm.fine_steps : prm.coarse_steps;
    trimg_cheb(bk, l, b, x, d, r, steps, prm.ratio);  // pre-smoothing from zero
    bk.residual(l, b, x, r);
    bk.restrict_to(l, r, bk.vec(l + 1, TRIMG_B));
    trimg_vcycle(bk, l + 1, prm);
    bk.prolong_add(l, bk.vec(l + 1, TRIMG_X), x);
    bk.residual(l, b, x, r);  // post-smoothing: x += q(A)(b - A x), the adjoint of the pre-smoother
    trimg_cheb(bk, l, r, e, d, bk.vec(l, TRIMG_B), steps, prm.ratio);
    bk.add(l, e, x);
}

}  // namespace tmx
