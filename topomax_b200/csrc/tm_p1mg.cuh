// Geometric multigrid for the Helmholtz filter operator  A_f = eps^2 K1 + M1  on the P1 vertex
// grid (reference: FEM_src/filter.py:27-41; there a MUMPS factorisation per call).
//
// Jacobi/Chebyshev iterations need O(eps/h) steps; on the large meshes (and on sharded runs,
// where every step costs a halo exchange) the filter then dominates the optimisation step.
// P1 spaces on uniformly coarsened right-diagonal meshes are nested and the Galerkin product
// P^T A P of a 7-point operator is again 7-point (centre, E, W, N, S, NE, SW), so coarse levels
// store one 7-coefficient stencil per vertex, built ONCE per engine (the operator does not depend
// on the design).  Level 0 keeps the analytic element form (tm_p1.cuh).  Odd cell counts coarsen
// with ceil(n/2) exactly like the P2 hierarchy: overhanging coarse vertices interpolate only to
// fine vertices that exist.  The levels share the P2 hierarchy's sizes and row-strip partition.
#pragma once

#include "tm_p1.cuh"

namespace tmx {

// stencil slot k <-> vertex offset (dx, dy): C, E, W, N, S, NE, SW
__host__ __device__ __forceinline__ void p1_offset(int k, int& dx, int& dy) {
    const int ox[7] = {0, 1, -1, 0, 0, 1, -1};
    const int oy[7] = {0, 0, 0, 1, -1, 1, -1};
    dx = ox[k];
    dy = oy[k];
}
__host__ __device__ __forceinline__ int p1_slot(int dx, int dy) {
    if (dx == 0 && dy == 0) return 0;
    if (dx == 1 && dy == 0) return 1;
    if (dx == -1 && dy == 0) return 2;
    if (dx == 0 && dy == 1) return 3;
    if (dx == 0 && dy == -1) return 4;
    if (dx == 1 && dy == 1) return 5;
    if (dx == -1 && dy == -1) return 6;
    return -1;
}
// linear interpolation weight of the coarse vertex at fine offset (dx, dy) from it
__host__ __device__ __forceinline__ double p1_interp_weight(int dx, int dy) {
    const int k = p1_slot(dx, dy);
    return k < 0 ? 0.0 : (k == 0 ? 1.0 : 0.5);
}

// one level as the kernels see it (g describes the LOCAL vertex grid; S == nullptr: analytic)
template <typename T>
struct P1Level {
    P1Geom g;
    double alpha, beta;
    const T* S;  // 7 planes of (g.nx+1)*(g.ny+1)
};

// entry (row vertex (ix,iy), slot k) of the level operator; vertices are LOCAL indices
template <typename T>
__device__ __forceinline__ double p1_level_entry(const P1Level<T>& L, int ix, int iy, int k) {
    int dx, dy;
    p1_offset(k, dx, dy);
    const int jx = ix + dx, jy = iy + dy;
    if (jx < 0 || jx > L.g.nx || jy < 0 || jy > L.g.ny) return 0.0;
    const size_t W1 = L.g.nx + 1;
    if (L.S) return (double)L.S[(size_t)k * W1 * (L.g.ny + 1) + (size_t)iy * W1 + ix];
    const size_t target = (size_t)jy * W1 + jx;
    return p1_row_apply(L.g, L.alpha, L.beta, ix, iy, [&](size_t j) { return j == target ? 1.0 : 0.0; });
}

// row (ix,iy) of the level operator applied to the field f(vertex index)
template <typename T, class F>
__device__ __forceinline__ double p1_level_row(const P1Level<T>& L, int ix, int iy, F&& f) {
    if (!L.S) return p1_row_apply(L.g, L.alpha, L.beta, ix, iy, f);
    const size_t W1 = L.g.nx + 1, n = W1 * (L.g.ny + 1), v = (size_t)iy * W1 + ix;
    double acc = (double)L.S[v] * f(v);
    if (ix < L.g.nx) acc += (double)L.S[n + v] * f(v + 1);
    if (ix > 0) acc += (double)L.S[2 * n + v] * f(v - 1);
    if (iy < L.g.ny) acc += (double)L.S[3 * n + v] * f(v + W1);
    if (iy > 0) acc += (double)L.S[4 * n + v] * f(v - W1);
    if (ix < L.g.nx && iy < L.g.ny) acc += (double)L.S[5 * n + v] * f(v + W1 + 1);
    if (ix > 0 && iy > 0) acc += (double)L.S[6 * n + v] * f(v - W1 - 1);
    return acc;
}

enum { P1EP_PLAIN = 0, P1EP_DOT = 1, P1EP_RESID = 2, P1EP_CHEB = 3 };

// y = A x (plain / with x.Ax),  y = b - A x,  or one Chebyshev-Jacobi step (d, y = x + d)
template <typename T, int EP>
__global__ void p1mg_apply_kernel(const P1Level<T> L, const T* __restrict__ x, T* __restrict__ y,
                                  const T* __restrict__ b, const T* __restrict__ dinv, T* __restrict__ d,
                                  double c1, double c2, ReduceScratch rs, double* dot_out) {
    const int ix = blockIdx.x * blockDim.x + threadIdx.x;
    const int iy = blockIdx.y * blockDim.y + threadIdx.y;
    double dot = 0.0;
    if (ix <= L.g.nx && iy <= L.g.ny && L.g.owns_row(iy)) {
        const size_t v = (size_t)iy * (L.g.nx + 1) + ix;
        const double ax = p1_level_row(L, ix, iy, [&](size_t j) { return (double)x[j]; });
        if (EP == P1EP_PLAIN || EP == P1EP_DOT) {
            y[v] = (T)ax;
            if (EP == P1EP_DOT) dot = (double)x[v] * ax;
        } else {
            const double r = (double)b[v] - ax;
            if (EP == P1EP_RESID) {
                y[v] = (T)r;
            } else {
                const double dn = (c1 != 0.0 ? c1 * (double)d[v] : 0.0) + c2 * (double)dinv[v] * r;
                d[v] = (T)dn;
                y[v] = (T)((double)x[v] + dn);
            }
        }
    }
    if (EP == P1EP_DOT) {
        double val[1] = {dot};
        double* const outs[1] = {dot_out};
        grid_reduce<1>(val, rs, outs);
    }
}

template <typename T>
__global__ void p1mg_diag_kernel(const P1Level<T> L, T* __restrict__ dinv) {
    const int ix = blockIdx.x * blockDim.x + threadIdx.x;
    const int iy = blockIdx.y * blockDim.y + threadIdx.y;
    if (ix > L.g.nx || iy > L.g.ny) return;
    const double dd = p1_level_entry(L, ix, iy, 0);
    dinv[(size_t)iy * (L.g.nx + 1) + ix] = (T)(dd != 0.0 ? 1.0 / dd : 1.0);
}

// Galerkin coarse stencils  S_c = P^T A_f P  for the owned coarse rows.
// f_off / c_off: global vertex row of local row 0 on the fine / coarse level; nxf_g, nyf_g: global
// fine cell counts (fine vertices beyond them do not exist).
template <typename T>
__global__ void p1mg_coarsen_kernel(const P1Level<T> F, int nyf_g, int f_off, int nxc, int nyc_loc, int c_off,
                                    int own0, int own1, T* __restrict__ Sc) {
    const int I = blockIdx.x * blockDim.x + threadIdx.x;
    const int J = blockIdx.y * blockDim.y + threadIdx.y;
    if (I > nxc || J > nyc_loc || J < own0 || J >= own1) return;
    const int Jg = J + c_off;
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int s1 = 0; s1 < 7; ++s1) {  // fine vertex i in the support of coarse (I, Jg)
        int d1x, d1y;
        p1_offset(s1, d1x, d1y);
        const int ixg = 2 * I + d1x, iyg = 2 * Jg + d1y;
        if (ixg < 0 || ixg > F.g.nx || iyg < 0 || iyg > nyf_g) continue;
        const int iyl = iyg - f_off;
        if (iyl < 0 || iyl > F.g.ny) continue;
        const double wi = s1 == 0 ? 1.0 : 0.5;
        for (int k = 0; k < 7; ++k) {  // fine neighbour j of i
            int ex, ey;
            p1_offset(k, ex, ey);
            const int jxg = ixg + ex, jyg = iyg + ey;
            if (jxg < 0 || jxg > F.g.nx || jyg < 0 || jyg > nyf_g) continue;
            const double a = p1_level_entry(F, ixg, iyl, k);
            if (a == 0.0) continue;
            // coarse vertices that j interpolates from: J' = I + off_K with weight w(j - 2J')
            for (int K = 0; K < 7; ++K) {
                int Kx, Ky;
                p1_offset(K, Kx, Ky);
                const double wj = p1_interp_weight(jxg - 2 * (I + Kx), jyg - 2 * (Jg + Ky));
                if (wj != 0.0) acc[K] += wi * a * wj;
            }
        }
    }
    const size_t W1 = nxc + 1, n = W1 * (nyc_loc + 1), v = (size_t)J * W1 + I;
    for (int K = 0; K < 7; ++K) Sc[K * n + v] = (T)acc[K];
}

// bc = P^T r on the owned coarse rows
template <typename T>
__global__ void p1mg_restrict_kernel(const P1Geom f, int nyf_g, const P1Geom c, const T* __restrict__ r,
                                     T* __restrict__ bc) {
    const int I = blockIdx.x * blockDim.x + threadIdx.x;
    const int J = blockIdx.y * blockDim.y + threadIdx.y;
    if (I > c.nx || J > c.ny || !c.owns_row(J)) return;
    const int Jg = J + c.iy_off;
    double acc = 0.0;
    for (int s = 0; s < 7; ++s) {
        int dx, dy;
        p1_offset(s, dx, dy);
        const int ixg = 2 * I + dx, iyg = 2 * Jg + dy;
        if (ixg < 0 || ixg > f.nx || iyg < 0 || iyg > nyf_g) continue;
        const int iyl = iyg - f.iy_off;
        if (iyl < 0 || iyl > f.ny) continue;
        acc += (s == 0 ? 1.0 : 0.5) * (double)r[(size_t)iyl * (f.nx + 1) + ixg];
    }
    bc[(size_t)J * (c.nx + 1) + I] = (T)acc;
}

// x += P xc on the owned fine rows
template <typename T>
__global__ void p1mg_prolong_add_kernel(const P1Geom f, const P1Geom c, const T* __restrict__ xc,
                                        T* __restrict__ x) {
    const int ix = blockIdx.x * blockDim.x + threadIdx.x;
    const int iy = blockIdx.y * blockDim.y + threadIdx.y;
    if (ix > f.nx || iy > f.ny || !f.owns_row(iy)) return;
    const int iyg = iy + f.iy_off;
    const size_t Wc = c.nx + 1;
    // both even: the coarse vertex itself; otherwise the two ends of the coarse edge it bisects
    const int I0 = ix >> 1, J0 = (iyg >> 1) - c.iy_off;
    const int I1 = (ix + 1) >> 1, J1 = ((iyg + 1) >> 1) - c.iy_off;
    const double v = 0.5 * ((double)xc[(size_t)J0 * Wc + I0] + (double)xc[(size_t)J1 * Wc + I1]);
    const size_t n = (size_t)iy * (f.nx + 1) + ix;
    x[n] = (T)((double)x[n] + v);
}

// coarsest level (<= 3x3 vertices): dense Cholesky factor from the stencils, and the solve
constexpr int kP1CoarseMax = 25;

template <typename T>
__global__ void p1mg_coarse_factor_kernel(const P1Level<T> L, double* __restrict__ A) {
    const int W1 = L.g.nx + 1, n = W1 * (L.g.ny + 1);
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) A[e] = 0.0;
    __syncthreads();
    for (int v = threadIdx.x; v < n; v += blockDim.x) {
        const int iy = v / W1, ix = v - iy * W1;
        for (int k = 0; k < 7; ++k) {
            int dx, dy;
            p1_offset(k, dx, dy);
            const int jx = ix + dx, jy = iy + dy;
            if (jx < 0 || jx > L.g.nx || jy < 0 || jy > L.g.ny) continue;
            A[v * n + jy * W1 + jx] = p1_level_entry(L, ix, iy, k);
        }
    }
    __syncthreads();
    for (int k = 0; k < n; ++k) {
        if (threadIdx.x == 0) A[k * n + k] = sqrt(A[k * n + k]);
        __syncthreads();
        const double dk = A[k * n + k];
        for (int i = k + 1 + threadIdx.x; i < n; i += blockDim.x) A[i * n + k] /= dk;
        __syncthreads();
        const int m = n - k - 1;
        for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
            const int i = k + 1 + e / m, j = k + 1 + e % m;
            if (j <= i) A[i * n + j] -= A[i * n + k] * A[j * n + k];
        }
        __syncthreads();
    }
}

template <typename T>
__global__ void p1mg_coarse_solve_kernel(int n, const double* __restrict__ Lm, const T* __restrict__ b,
                                         T* __restrict__ x) {
    __shared__ double y[kP1CoarseMax];
    for (int i = threadIdx.x; i < n; i += blockDim.x) y[i] = (double)b[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < n; ++k) {
            y[k] /= Lm[k * n + k];
            for (int i = k + 1; i < n; ++i) y[i] -= Lm[i * n + k] * y[k];
        }
        for (int k = n - 1; k >= 0; --k) {
            y[k] /= Lm[k * n + k];
            for (int i = 0; i < k; ++i) y[i] -= Lm[k * n + i] * y[k];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) x[i] = (T)y[i];
}

}  // namespace tmx
