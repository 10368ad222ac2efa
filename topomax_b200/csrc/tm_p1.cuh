// Kernels on the P1 vertex grid (density, filter, sensitivity, mirror descent) and the one-off
// load vector.  All are node-centric gathers over the <= 6 incident triangles: written once
// per node by one thread, fixed summation order, no atomics.
#pragma once

#include "tm_common.cuh"

namespace tmx {

// ---------------------------------------------------------------------------------------
// P1 grid description + element matrices of the two triangle orientations
// ---------------------------------------------------------------------------------------
struct P1Geom {
    int nx, ny;            // cells; vertices (nx+1) x (ny+1)
    double hx, hy;
    double Ke[2][3][3];    // |T| grad(lambda_i).grad(lambda_j), T_A then T_B
    double Me[3][3];       // |T|/12 (1 + delta_ij)
    // assembled interior rows of K1 and M1: centre, E, W, N, S, NE, SW (the "right" diagonal
    // couples NE/SW only); boundary vertices go through the element loop
    double sK[7], sM[7];
    // row-strip sharding: local vertex row 0 is global row iy_off; only rows in
    // [own_iy0, own_iy1) are written / summed by this rank (one GPU: 0, ny_global = ny, all rows)
    int iy_off, ny_global, own_iy0, own_iy1;
    __host__ __device__ bool owns_row(int iy) const { return iy >= own_iy0 && iy < own_iy1; }
};

// fills sK/sM by assembling the row of the centre vertex of a 2x2-cell patch
inline void p1_fill_interior_stencil(P1Geom& g);

// triangle-local vertex (0..2) of cell corner (ax,ay) in {0,1}^2, or -1
__host__ __device__ __forceinline__ int corner_to_tri_local(int type, int ax, int ay) {
    if (ax == 0 && ay == 0) return 0;
    if (ax == 1 && ay == 1) return 2;
    if (type == 0) return (ax == 1 && ay == 0) ? 1 : -1;
    return (ax == 0 && ay == 1) ? 1 : -1;
}
// vertex index of triangle-local vertex k of cell (cx,cy)
__host__ __device__ __forceinline__ size_t tri_vertex(int type, int k, int cx, int cy, int nx) {
    const int ax = (k == 2) ? 1 : ((k == 1 && type == 0) ? 1 : 0);
    const int ay = (k == 2) ? 1 : ((k == 1 && type == 1) ? 1 : 0);
    return (size_t)(cy + ay) * (nx + 1) + (cx + ax);
}

// row (ix,iy) of alpha K1 + beta M1 applied to the field f(vertex index)
template <class F>
__device__ __forceinline__ double p1_row_apply(const P1Geom& g, double alpha, double beta, int ix,
                                               int iy, F&& f) {
    if (ix > 0 && ix < g.nx && iy > 0 && iy < g.ny) {
        const size_t v = (size_t)iy * (g.nx + 1) + ix;
        const size_t W1 = g.nx + 1;
        double acc = (alpha * g.sK[0] + beta * g.sM[0]) * f(v);
        acc += (alpha * g.sK[1] + beta * g.sM[1]) * f(v + 1);
        acc += (alpha * g.sK[2] + beta * g.sM[2]) * f(v - 1);
        acc += (alpha * g.sK[3] + beta * g.sM[3]) * f(v + W1);
        acc += (alpha * g.sK[4] + beta * g.sM[4]) * f(v - W1);
        acc += (alpha * g.sK[5] + beta * g.sM[5]) * f(v + W1 + 1);
        acc += (alpha * g.sK[6] + beta * g.sM[6]) * f(v - W1 - 1);
        return acc;
    }
    double acc = 0.0;
    for (int cy = max(iy - 1, 0); cy <= min(iy, g.ny - 1); ++cy)
        for (int cx = max(ix - 1, 0); cx <= min(ix, g.nx - 1); ++cx) {
#pragma unroll
            for (int type = 0; type < 2; ++type) {
                const int kl = corner_to_tri_local(type, ix - cx, iy - cy);
                if (kl < 0) continue;
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    acc += (alpha * g.Ke[type][kl][k] + beta * g.Me[kl][k]) *
                           f(tri_vertex(type, k, cx, cy, g.nx));
            }
        }
    return acc;
}

inline void p1_fill_interior_stencil(P1Geom& g) {
    for (int k = 0; k < 7; ++k) g.sK[k] = g.sM[k] = 0.0;
    const int off[7][2] = {{0, 0}, {1, 0}, {-1, 0}, {0, 1}, {0, -1}, {1, 1}, {-1, -1}};
    // centre vertex (1,1) of a 2x2-cell patch
    for (int cy = 0; cy < 2; ++cy)
        for (int cx = 0; cx < 2; ++cx)
            for (int type = 0; type < 2; ++type) {
                const int kl = corner_to_tri_local(type, 1 - cx, 1 - cy);
                if (kl < 0) continue;
                for (int k = 0; k < 3; ++k) {
                    const int ax = (k == 2) ? 1 : ((k == 1 && type == 0) ? 1 : 0);
                    const int ay = (k == 2) ? 1 : ((k == 1 && type == 1) ? 1 : 0);
                    const int dx = cx + ax - 1, dy = cy + ay - 1;
                    for (int o = 0; o < 7; ++o)
                        if (off[o][0] == dx && off[o][1] == dy) {
                            g.sK[o] += g.Ke[type][kl][k];
                            g.sM[o] += g.Me[kl][k];
                        }
                }
            }
}

// y = (alpha K1 + beta M1) x    [optionally the grid-wide x . y]
// Filter operator A_f = eps^2 K1 + M1 (reference: FEM_src/filter.py:27-33); alpha=0, beta=1
// gives the consistent-mass product M1 rho of the filter's right-hand side (:35-36).
template <typename T, bool DOT>
__global__ void p1_apply_kernel(const P1Geom g, double alpha, double beta, const T* __restrict__ x,
                                T* __restrict__ y, ReduceScratch rs, double* dot_out) {
    const int ix = blockIdx.x * blockDim.x + threadIdx.x;
    const int iy = blockIdx.y * blockDim.y + threadIdx.y;
    double dot = 0.0;
    if (ix <= g.nx && iy <= g.ny && g.owns_row(iy)) {
        const double acc = p1_row_apply(g, alpha, beta, ix, iy, [&](size_t j) { return (double)x[j]; });
        const size_t v = (size_t)iy * (g.nx + 1) + ix;
        y[v] = (T)acc;
        if (DOT) dot = (double)x[v] * acc;
    }
    if (DOT) {
        double val[1] = {dot};
        double* const outs[1] = {dot_out};
        grid_reduce<1>(val, rs, outs);
    }
}

// inverse diagonal of alpha K1 + beta M1
template <typename T>
__global__ void p1_diag_kernel(const P1Geom g, double alpha, double beta, T* __restrict__ dinv) {
    const int ix = blockIdx.x * blockDim.x + threadIdx.x;
    const int iy = blockIdx.y * blockDim.y + threadIdx.y;
    if (ix > g.nx || iy > g.ny) return;
    double acc = 0.0;
    for (int cy = max(iy - 1, 0); cy <= min(iy, g.ny - 1); ++cy)
        for (int cx = max(ix - 1, 0); cx <= min(ix, g.nx - 1); ++cx)
            for (int type = 0; type < 2; ++type) {
                const int kl = corner_to_tri_local(type, ix - cx, iy - cy);
                if (kl < 0) continue;
                acc += alpha * g.Ke[type][kl][kl] + beta * g.Me[kl][kl];
            }
    dinv[(size_t)iy * (g.nx + 1) + ix] = (T)(1.0 / acc);
}

// nodal quadrature weight w_i = (M1 . 1)_i = |T|/3 * (#incident triangles)
// (reference: FEM_src/solver.py:81-84, integrate = assemble(f*dx) of the P1 interpolant)
__device__ __forceinline__ double p1_weight(const P1Geom& g, int ix, int iy_local) {
    int cnt = 0;
    const int iy = iy_local + g.iy_off;  // global vertex row decides the boundary weights
    for (int cy = max(iy - 1, 0); cy <= min(iy, g.ny_global - 1); ++cy)
        for (int cx = max(ix - 1, 0); cx <= min(ix, g.nx - 1); ++cx)
            for (int type = 0; type < 2; ++type)
                if (corner_to_tri_local(type, ix - cx, iy - cy) >= 0) ++cnt;
    return cnt * (g.hx * g.hy / 6.0);
}

__device__ __forceinline__ double expit_d(double x) { return 1.0 / (1.0 + exp(-x)); }

// Volume-projection evaluation (reference: src/solver.py:158-162):
//   out[0] = sum_i w_i expit(half_i + c),  out[1] = sum_i w_i expit'(half_i + c)
// c_dev != nullptr: the shift is read from device memory (the device-resident Newton iteration)
template <typename T>
__global__ void md_volume_kernel(const P1Geom g, const T* __restrict__ half, double c,
                                 ReduceScratch rs, double* out, const double* __restrict__ c_dev = nullptr) {
    if (c_dev) c = *c_dev;
    double val[2] = {0.0, 0.0};
    const size_t n1 = (size_t)(g.nx + 1) * (g.ny + 1);
    for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < n1;
         v += (size_t)gridDim.x * blockDim.x) {
        const int iy = (int)(v / (g.nx + 1)), ix = (int)(v - (size_t)iy * (g.nx + 1));
        if (!g.owns_row(iy)) continue;
        const double w = p1_weight(g, ix, iy);
        const double s = expit_d((double)half[v] + c);
        val[0] += w * s;
        val[1] += w * (s * (1.0 - s));
    }
    double* const outs[2] = {out, out + 1};
    grid_reduce<2>(val, rs, outs);
}

// One Newton iterate of the volume projection on the device (reference: src/solver.py:166-174 =
// scipy.optimize.newton(error, 0, fprime, tol=1e-12): p = p0 - f(p0)/f'(p0), converged when
// |p - p0| <= tol; f(p0) == 0 returns p0; f'(p0) == 0 fails).  st = {c, done, iterations, status}
// (status 1 converged, 2 zero derivative); vol[0] = sum w expit(half + c), vol[1] = its derivative.
// A finished iteration turns the following updates into no-ops, so a batch can be enqueued blind.
__global__ void md_newton_update_kernel(const double* __restrict__ vol, double volume, double tol,
                                        double* __restrict__ st) {
    if (threadIdx.x != 0 || blockIdx.x != 0 || st[1] != 0.0) return;
    const double c = st[0];
    const double f = vol[0] - volume;
    if (f == 0.0) {
        st[1] = 1.0;
        st[3] = 1.0;
        return;
    }
    const double fder = vol[1];
    if (fder == 0.0) {
        st[1] = 1.0;
        st[3] = 2.0;
        return;
    }
    const double p = c - f / fder;
    st[2] += 1.0;
    st[0] = p;
    if (fabs(p - c) <= tol) {
        st[1] = 1.0;
        st[3] = 1.0;
    }
}

// psi = half + c; rho = expit(psi); out[0] = sum w (rho - expit(psi_prev))^2; out[1] = sum w rho
// (reference: src/solver.py:262,286-288)
template <typename T>
__global__ void md_apply_kernel(const P1Geom g, const T* __restrict__ half, double c,
                                const T* __restrict__ psi_prev, T* __restrict__ psi,
                                T* __restrict__ rho, ReduceScratch rs, double* out) {
    double val[2] = {0.0, 0.0};
    const size_t n1 = (size_t)(g.nx + 1) * (g.ny + 1);
    for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < n1;
         v += (size_t)gridDim.x * blockDim.x) {
        const int iy = (int)(v / (g.nx + 1)), ix = (int)(v - (size_t)iy * (g.nx + 1));
        if (!g.owns_row(iy)) continue;
        const double w = p1_weight(g, ix, iy);
        const double prev = expit_d((double)psi_prev[v]);
        const double p = (double)half[v] + c;
        const double r = expit_d(p);
        psi[v] = (T)p;
        rho[v] = (T)r;
        val[0] += w * (r - prev) * (r - prev);
        val[1] += w * r;
    }
    double* const outs[2] = {out, out + 1};
    grid_reduce<2>(val, rs, outs);
}

// out[0] = sum_i w_i values_i   (FEMSolver.integrate, reference: FEM_src/solver.py:81-84)
template <typename T>
__global__ void p1_integrate_kernel(const P1Geom g, const T* __restrict__ values, ReduceScratch rs,
                                    double* out) {
    double val[1] = {0.0};
    const size_t n1 = (size_t)(g.nx + 1) * (g.ny + 1);
    for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < n1;
         v += (size_t)gridDim.x * blockDim.x) {
        const int iy = (int)(v / (g.nx + 1)), ix = (int)(v - (size_t)iy * (g.nx + 1));
        if (!g.owns_row(iy)) continue;
        val[0] += p1_weight(g, ix, iy) * (double)values[v];
    }
    double* const outs[1] = {out};
    grid_reduce<1>(val, rs, outs);
}

// ---------------------------------------------------------------------------------------
// sensitivity right-hand side  b^g_i = int -r'(xi_h) (lambda (div u)^2 + 2 mu eps:eps) phi_i
// (reference: FEM_src/elasisity_problem.py:146-150), gathered per vertex
// ---------------------------------------------------------------------------------------
// GENERAL = false: p = 3 closed form; true: any exponent (tri_sensitivity_general)
template <typename T, bool GENERAL>
__global__ void sens_rhs_kernel(const LevelGeom<T> g, const PenaltySpec ps, int own_iy0, int own_iy1,
                                const T* __restrict__ u, T* __restrict__ out) {
    const int ix = blockIdx.x * blockDim.x + threadIdx.x;
    const int iy = blockIdx.y * blockDim.y + threadIdx.y;
    if (ix > g.nx || iy > g.ny || iy < own_iy0 || iy >= own_iy1) return;
    double acc = 0.0;
    for (int cy = max(iy - 1, 0); cy <= min(iy, g.ny - 1); ++cy)
        for (int cx = max(ix - 1, 0); cx <= min(ix, g.nx - 1); ++cx) {
#pragma unroll
            for (int type = 0; type < 2; ++type) {
                const int kl = corner_to_tri_local(type, ix - cx, iy - cy);
                if (kl < 0) continue;
                T U[3][2], M[3][2], xi[3], e[3][3], gl[3];
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    const int q = tri_local_to_cell(type, k);
                    const size_t n = (size_t)(2 * cy + q / 3) * g.Lx + (2 * cx + q % 3);
                    const T a = u[2 * n], b = u[2 * n + 1];
                    if (k < 3) {
                        U[k][0] = a; U[k][1] = b;
                    } else {
                        M[k - 3][0] = a; M[k - 3][1] = b;
                    }
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) xi[k] = g.xi[tri_vertex(type, k, cx, cy, g.nx)];
                if (type == 0)
                    tri_vertex_strains<T, false>(U, M, g.mat.kappa, e);
                else
                    tri_vertex_strains<T, true>(U, M, g.mat.kappa, e);
                if (GENERAL)
                    tri_sensitivity_general<T>(e, xi, g.simp_min, g.mat, ps, gl);
                else
                    tri_sensitivity<T>(e, xi, g.simp_min, g.mat, gl);
                acc += (double)gl[kl];
            }
        }
    out[(size_t)iy * (g.nx + 1) + ix] = (T)acc;
}

// ---------------------------------------------------------------------------------------
// load vector  b = int f_h . v dx + int t_h . v ds,  f_h / t_h the P2 nodal interpolants of
// the reference's indicator expressions (reference: FEM_src/elasisity_problem.py:20-73,
// :120-124; dolfin interpolates a UserExpression without degree into P2: SURVEY App. A.5).
// Node coordinates and the inside tests follow the reference's floating-point recipe exactly
// (round-to-nearest ops, no FMA contraction): which nodes are "inside" decides the load.
// ---------------------------------------------------------------------------------------
struct LoadSpec {
    int nx, ny;        // global cell counts
    int j_off, Ly_loc; // local lattice rows [j_off, j_off + Ly_loc) are produced
    double W, H;
    int has_force;
    double fcx, fcy, frad, fx, fy;
    int ntractions;
    int tside[8];  // 0 Left 1 Right 2 Top 3 Bottom
    double tlo[8], thi[8], tx[8], ty[8];
    double M2[6][6];  // P2 mass matrix of one triangle (same for both orientations)
};

__device__ __forceinline__ double lattice_coord(int i, double extent, int ncell) {
    // dolfin RectangleMesh: vertex = (k*extent)/n ; P2 midpoint = 0.5*a + 0.5*b
    if ((i & 1) == 0) return __ddiv_rn(__dmul_rn((double)(i >> 1), extent), (double)ncell);
    const double a = __ddiv_rn(__dmul_rn((double)((i - 1) >> 1), extent), (double)ncell);
    const double b = __ddiv_rn(__dmul_rn((double)((i + 1) >> 1), extent), (double)ncell);
    return __dadd_rn(__dmul_rn(0.5, a), __dmul_rn(0.5, b));
}

__device__ __forceinline__ bool inside_disc(const LoadSpec& s, int i, int j) {
    const double dx = __dsub_rn(lattice_coord(i, s.W, s.nx), s.fcx);
    const double dy = __dsub_rn(lattice_coord(j, s.H, s.ny), s.fcy);
    const double dist = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    return dist < s.frad;
}

// Nodal traction value of the boundary node at lattice position p along side `side`: the sum of
// EVERY traction whose own side passes through the node and whose window contains it
// (TractionExpression.eval, FEM_src/elasisity_problem.py:47-70, is evaluated at the node, whatever
// facet is being integrated).  Away from the corners only tractions of `side` can match; a CORNER
// node also picks up a traction of the adjacent side whose window reaches the corner, and through
// the P2 interpolant on the corner edge of `side` that value loads this side too.
__device__ __forceinline__ void traction_nodal(const LoadSpec& s, int side, int p, double& tx,
                                               double& ty) {
    tx = 0.0;
    ty = 0.0;
    const int Lx = 2 * s.nx + 1, Ly = 2 * s.ny + 1;
    const int i = side == 0 ? 0 : (side == 1 ? Lx - 1 : p);
    const int j = side == 2 ? Ly - 1 : (side == 3 ? 0 : p);
    for (int t = 0; t < s.ntractions; ++t) {
        const int ts = s.tside[t];
        const bool on = (ts == 0 && i == 0) || (ts == 1 && i == Lx - 1) || (ts == 2 && j == Ly - 1) ||
                        (ts == 3 && j == 0);
        if (!on) continue;
        const double c = ts <= 1 ? lattice_coord(j, s.H, s.ny) : lattice_coord(i, s.W, s.nx);
        // df.between(c, (lo, hi)) with DOLFIN_EPS = 3e-16
        if (c >= __dsub_rn(s.tlo[t], 3.0e-16) && c <= __dadd_rn(s.thi[t], 3.0e-16)) {
            tx += s.tx[t];
            ty += s.ty[t];
        }
    }
}

template <typename T>
__global__ void load_vector_kernel(const LoadSpec s, T* __restrict__ b) {
    const int Lx = 2 * s.nx + 1, Ly = 2 * s.ny + 1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int jl = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= Lx || jl >= s.Ly_loc) return;
    const int j = jl + s.j_off;  // global lattice row
    double b0 = 0.0, b1 = 0.0;
    if (s.has_force) {
        double acc = 0.0;
        const int cx_lo = i > 0 ? (i - 1) >> 1 : 0, cx_hi = min(s.nx - 1, i >> 1);
        const int cy_lo = j > 0 ? (j - 1) >> 1 : 0, cy_hi = min(s.ny - 1, j >> 1);
        for (int cy = cy_lo; cy <= cy_hi; ++cy)
            for (int cx = cx_lo; cx <= cx_hi; ++cx) {
                const int q = 3 * (j - 2 * cy) + (i - 2 * cx);
                for (int type = 0; type < 2; ++type) {
                    const int kl = cell_to_tri_local(type, q);
                    if (kl < 0) continue;
                    for (int k = 0; k < 6; ++k) {
                        const int qq = tri_local_to_cell(type, k);
                        if (inside_disc(s, 2 * cx + qq % 3, 2 * cy + qq / 3)) acc += s.M2[kl][k];
                    }
                }
            }
        b0 += s.fx * acc;
        b1 += s.fy * acc;
    }
    // boundary edges: exact 1-D P2 mass  h/30 [[4,2,-1],[2,16,2],[-1,2,4]] on (v0, mid, v1)
    for (int side = 0; side < 4; ++side) {
        bool on;
        int p, np;
        double hlen;
        if (side == 0) { on = (i == 0); p = j; np = Ly; hlen = s.H / s.ny; }
        else if (side == 1) { on = (i == Lx - 1); p = j; np = Ly; hlen = s.H / s.ny; }
        else if (side == 2) { on = (j == Ly - 1); p = i; np = Lx; hlen = s.W / s.nx; }
        else { on = (j == 0); p = i; np = Lx; hlen = s.W / s.nx; }
        if (!on || s.ntractions == 0) continue;
        double a0 = 0.0, a1 = 0.0;
        const double c = hlen / 30.0;
        if (p & 1) {  // midpoint of edge (p-1, p, p+1)
            double tx, ty;
            traction_nodal(s, side, p - 1, tx, ty); a0 += 2 * c * tx; a1 += 2 * c * ty;
            traction_nodal(s, side, p, tx, ty);     a0 += 16 * c * tx; a1 += 16 * c * ty;
            traction_nodal(s, side, p + 1, tx, ty); a0 += 2 * c * tx; a1 += 2 * c * ty;
        } else {
            double tx, ty;
            if (p >= 2) {  // edge below/left: this node is its v1
                traction_nodal(s, side, p - 2, tx, ty); a0 += -c * tx; a1 += -c * ty;
                traction_nodal(s, side, p - 1, tx, ty); a0 += 2 * c * tx; a1 += 2 * c * ty;
                traction_nodal(s, side, p, tx, ty);     a0 += 4 * c * tx; a1 += 4 * c * ty;
            }
            if (p + 2 < np) {  // edge above/right: this node is its v0
                traction_nodal(s, side, p, tx, ty);     a0 += 4 * c * tx; a1 += 4 * c * ty;
                traction_nodal(s, side, p + 1, tx, ty); a0 += 2 * c * tx; a1 += 2 * c * ty;
                traction_nodal(s, side, p + 2, tx, ty); a0 += -c * tx; a1 += -c * ty;
            }
        }
        b0 += a0;
        b1 += a1;
    }
    const size_t n = (size_t)jl * Lx + i;
    b[2 * n] = (T)b0;
    b[2 * n + 1] = (T)b1;
}

// ---------------------------------------------------------------------------------------
// point evaluation of a P1 (COMP = 1) or vector-P2 (COMP = 2) field on a regular grid of sample
// points  (x, y) = (x0 + sx dx, y0 + sy dy): the role of  f(x, y)  inside the double loop of the
// reference's sample_function (FEM_src/utils.py:112-162), which is what plot.py consumes.
// The containing cell is found by index arithmetic; s >= t selects T_A = (v0, v1, v3), else
// T_B = (v0, v2, v3) ("right" diagonal); the fields are continuous, so points on cell or
// triangle boundaries get the same value from either side (to rounding).
// out[sy][sx][COMP], one thread per sample, coalesced stores.
// ---------------------------------------------------------------------------------------
template <typename T, int COMP>
__global__ void sample_field_kernel(int nx, int ny, double hx, double hy, const T* __restrict__ f, int nsx,
                                    int nsy, double x0, double dx, double y0, double dy, T* __restrict__ out) {
    const int sx = blockIdx.x * blockDim.x + threadIdx.x;
    const int sy = blockIdx.y * blockDim.y + threadIdx.y;
    if (sx >= nsx || sy >= nsy) return;
    const double x = x0 + sx * dx, y = y0 + sy * dy;
    int cx = (int)floor(x / hx), cy = (int)floor(y / hy);
    cx = min(max(cx, 0), nx - 1);
    cy = min(max(cy, 0), ny - 1);
    const double s = x / hx - cx, t = y / hy - cy;  // local coordinates in the cell
    const bool lower = s >= t;                       // T_A below the diagonal v0 -> v3
    // barycentrics of (v0, vm, v3) with vm = v1 (T_A) or v2 (T_B)
    const double l0 = lower ? 1.0 - s : 1.0 - t;
    const double lm = lower ? s - t : t - s;
    const double l3 = lower ? t : s;
    const size_t o = ((size_t)sy * nsx + sx) * COMP;
    if (COMP == 1) {
        const size_t W1 = nx + 1, v0 = (size_t)cy * W1 + cx;
        const size_t vm = lower ? v0 + 1 : v0 + W1;
        out[o] = (T)(l0 * (double)f[v0] + lm * (double)f[vm] + l3 * (double)f[v0 + W1 + 1]);
    } else {
        const size_t Lx = 2 * (size_t)nx + 1;
        const int i0 = 2 * cx, j0 = 2 * cy;
        // lattice offsets of (v0, vm, v3, mid(v0,vm), mid(vm,v3), mid(v0,v3))
        const int mi = lower ? 2 : 0, mj = lower ? 0 : 2;
        const int di[6] = {0, mi, 2, mi / 2, (mi + 2) / 2, 1};
        const int dj[6] = {0, mj, 2, mj / 2, (mj + 2) / 2, 1};
        const double phi[6] = {l0 * (2.0 * l0 - 1.0), lm * (2.0 * lm - 1.0), l3 * (2.0 * l3 - 1.0),
                               4.0 * l0 * lm, 4.0 * lm * l3, 4.0 * l0 * l3};
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const size_t n = (size_t)(j0 + dj[k]) * Lx + (i0 + di[k]);
            a0 += phi[k] * (double)f[2 * n];
            a1 += phi[k] * (double)f[2 * n + 1];
        }
        out[o] = (T)a0;
        out[o + 1] = (T)a1;
    }
}

}  // namespace tmx
