// The tail of the multigrid hierarchy (levels of a few thousand lattice nodes and below) as ONE
// thread-block-cluster kernel per V-cycle.
//
// On those levels a launch does microseconds of work: the march-and-shuffle operator kernel is a
// serial latency chain (loads -> ~1500 dependent fp64 flops on one warp per SM sub-partition ->
// epilogue loads) plus a kernel boundary, ~5 us per launch, ~9 launches per level and V-cycle.
// Here the same Galerkin operators are assembled ONCE per hierarchy set-up into node stencils
// (5x5 lattice neighbourhood, 2x2 blocks: row n of A as 100 coefficients, SoA), so an operator
// application is one independent gather per node - no marching, no shuffles - and the phases of
// the V-cycle (Chebyshev steps, residual, restriction, coarsest solve, prolongation) are
// separated by the hardware cluster barrier instead of kernel boundaries.  The stencil
// coefficients of a node are always read by the same thread, so they stay in that SM's L1
// across phases (ld.global.nc); lattice vectors written by other CTAs are read through L2
// (ld.global.cg) after the barrier's release/acquire.
//
// The arithmetic is the smoother / transfer / coarse-solve sequence of Engine::vcycle_body on the
// assembled form of the same operators: results agree with the launch-per-phase path to
// round-off (summation order differs), which tests/test_gpu_parity.py checks.
#pragma once

#include <cooperative_groups.h>

#include "tm_common.cuh"
#include "tm_element.cuh"
#include "tm_mg.cuh"
#include "tm_tables.h"

namespace tmx {

namespace cg = cooperative_groups;

constexpr int kTailMaxLevels = 8;
constexpr int kTailThreads = 256;
constexpr int kTailMaxDegree = 8;

template <typename T>
struct TailLevel {
    LevelGeom<T> g;  // replicated geometry: j_off = 0, every row owned
    int n;           // lattice nodes Lx * Ly
    const T* S;      // [100][n]: S[4*nb + 2*c + c'][node], nb = 5*(dj+2) + (di+2)
    const T* dinv;   // [2n]
    T* b;            // right-hand side (first tail level: written by the launching V-cycle)
    T *xa, *xb, *d, *r;
    T c1[kTailMaxDegree], c2[kTailMaxDegree];  // Chebyshev step k: d = c1 d + c2 D^-1 r
};

template <typename T>
struct TailArgs {
    int nt;      // tail levels; the last one is the coarsest (dense inverse)
    int degree;  // Chebyshev-Jacobi steps before and after the coarse correction
    int nc;      // dofs of the coarsest level
    const double* Ainv;
    TransferTable tab;
    TailLevel<T> lv[kTailMaxLevels];
};

// Row (node, c) of the assembled level operator: each of the <= 4 cells around the node applies
// its element matrix to the unit vector of (node, c); by symmetry that column is the row.
// Dirichlet rows are identity rows, couplings to Dirichlet nodes are dropped (the operator
// kernels mask the same entries, tm_elast.cuh).
template <typename T>
__global__ void tail_assemble_kernel(const LevelGeom<T> g, T* __restrict__ S) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = g.Lx * g.Ly;
    if (t >= 2 * n) return;
    const int node = t >> 1, c = t & 1;
    const int i = node % g.Lx, j = node / g.Lx;
    T acc[25][2];
#pragma unroll
    for (int nb = 0; nb < 25; ++nb) acc[nb][0] = acc[nb][1] = T(0);
    if (g.fixed(i, j)) {
        acc[12][c] = T(1);
    } else {
        const int cx_lo = i > 0 ? (i - 1) >> 1 : 0, cx_hi = min(g.nx - 1, i >> 1);
        const int cy_lo = j > 0 ? (j - 1) >> 1 : 0, cy_hi = min(g.ny - 1, j >> 1);
        const size_t plane = (size_t)g.nx * g.ny;
        for (int cy = cy_lo; cy <= cy_hi; ++cy)
            for (int cx = cx_lo; cx <= cx_hi; ++cx) {
                const int qi = i - 2 * cx, qj = j - 2 * cy;
                T X[9][2], Y[9][2], wA[6], wB[6];
#pragma unroll
                for (int q = 0; q < 9; ++q) {
                    const bool me = (q == 3 * qj + qi);
                    X[q][0] = (me && c == 0) ? T(1) : T(0);
                    X[q][1] = (me && c == 1) ? T(1) : T(0);
                    Y[q][0] = Y[q][1] = T(0);
                }
                const size_t cidx = (size_t)cy * g.nx + cx;
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    wA[k] = g.W[k * plane + cidx];
                    wB[k] = g.W[(6 + k) * plane + cidx];
                }
                cell_apply<T>(X, wA, wB, g.mat, Y);
#pragma unroll
                for (int q = 0; q < 9; ++q) {
                    const int di = q % 3 - qi, dj = q / 3 - qj;
                    if (g.fixed(i + di, j + dj)) continue;
                    const int nb = 5 * (dj + 2) + (di + 2);
                    acc[nb][0] += Y[q][0];
                    acc[nb][1] += Y[q][1];
                }
            }
    }
    for (int nb = 0; nb < 25; ++nb) {
        S[(size_t)(4 * nb + 2 * c + 0) * n + node] = acc[nb][0];
        S[(size_t)(4 * nb + 2 * c + 1) * n + node] = acc[nb][1];
    }
}

template <typename T>
__device__ __forceinline__ typename Vec2<T>::type tail_ld2(const T* p, size_t node) {
    return __ldcg(reinterpret_cast<const typename Vec2<T>::type*>(p) + node);
}
template <typename T>
__device__ __forceinline__ void tail_st2(T* p, size_t node, T v0, T v1) {
    typename Vec2<T>::type v;
    v.x = v0;
    v.y = v1;
    __stcg(reinterpret_cast<typename Vec2<T>::type*>(p) + node, v);
}

// (A x)(node) from the stencil; ONFLY: x = c0 D^-1 b formed at every neighbour instead of read
// (the first Chebyshev step from a zero guess is never stored).  xo = x at the node itself.
template <typename T, bool ONFLY>
__device__ __forceinline__ void tail_row(const TailLevel<T>& L, int i, int j, int node,
                                         const T* cur, T c0, T& y0, T& y1, T& xo0, T& xo1) {
    using V2 = typename Vec2<T>::type;
    const int n = L.n, Lx = L.g.Lx, Ly = L.g.Ly;
    T a0 = T(0), a1 = T(0), e0 = T(0), e1 = T(0);  // two partial sums per component, fixed order
#pragma unroll
    for (int dj = -2; dj <= 2; ++dj)
#pragma unroll
        for (int di = -2; di <= 2; ++di) {
            const int nb = 5 * (dj + 2) + (di + 2);
            const int ii = i + di, jj = j + dj;
            const bool in = ii >= 0 && ii < Lx && jj >= 0 && jj < Ly;
            const int m = in ? jj * Lx + ii : node;  // outside: coefficient is exactly 0
            V2 xv;
            if (ONFLY) {
                const V2 bv = tail_ld2<T>(L.b, m);
                const V2 dv = __ldg(reinterpret_cast<const V2*>(L.dinv) + m);
                xv.x = c0 * dv.x * bv.x;
                xv.y = c0 * dv.y * bv.y;
            } else {
                xv = tail_ld2<T>(cur, m);
            }
            const T s00 = __ldg(L.S + (size_t)(4 * nb + 0) * n + node);
            const T s01 = __ldg(L.S + (size_t)(4 * nb + 1) * n + node);
            const T s10 = __ldg(L.S + (size_t)(4 * nb + 2) * n + node);
            const T s11 = __ldg(L.S + (size_t)(4 * nb + 3) * n + node);
            if (nb & 1) {
                e0 += s00 * xv.x + s01 * xv.y;
                e1 += s10 * xv.x + s11 * xv.y;
            } else {
                a0 += s00 * xv.x + s01 * xv.y;
                a1 += s10 * xv.x + s11 * xv.y;
            }
            if (nb == 12) {
                xo0 = xv.x;
                xo1 = xv.y;
            }
        }
    y0 = a0 + e0;
    y1 = a1 + e1;
}

// One smoothing / residual phase over the level.  CHEB: r = b - A x; d = c1 d + c2 D^-1 r;
// out = x + d.  !CHEB: out = b - A x (and, ONFLY, x itself is stored to xstore).
template <typename T, bool CHEB, bool ONFLY>
__device__ __forceinline__ void tail_phase(const TailLevel<T>& L, const T* cur, T* out, T* xstore,
                                           T c0, T c1, T c2, bool store_d, int gtid, int nthreads) {
    using V2 = typename Vec2<T>::type;
    for (int node = gtid; node < L.n; node += nthreads) {
        const int i = node % L.g.Lx, j = node / L.g.Lx;
        if (L.g.fixed(i, j)) {  // identity rows with zero data: everything stays zero
            tail_st2<T>(out, node, T(0), T(0));
            if (CHEB && store_d) tail_st2<T>(L.d, node, T(0), T(0));
            if (!CHEB && ONFLY) tail_st2<T>(xstore, node, T(0), T(0));
            continue;
        }
        T y0, y1, x0 = T(0), x1 = T(0);
        tail_row<T, ONFLY>(L, i, j, node, cur, c0, y0, y1, x0, x1);
        const V2 bv = tail_ld2<T>(L.b, node);
        const T r0 = bv.x - y0, r1 = bv.y - y1;
        if (!CHEB) {
            tail_st2<T>(out, node, r0, r1);
            if (ONFLY) tail_st2<T>(xstore, node, x0, x1);
        } else {
            const V2 dv = __ldg(reinterpret_cast<const V2*>(L.dinv) + node);
            T d0 = c2 * dv.x * r0, d1 = c2 * dv.y * r1;
            if (ONFLY) {  // the direction of the un-stored first step is x itself
                d0 += c1 * x0;
                d1 += c1 * x1;
            } else if (c1 != T(0)) {
                const V2 dp = tail_ld2<T>(L.d, node);
                d0 += c1 * dp.x;
                d1 += c1 * dp.y;
            }
            if (store_d) tail_st2<T>(L.d, node, d0, d1);
            tail_st2<T>(out, node, x0 + d0, x1 + d1);
        }
    }
}

// z_tail = V(b_tail): launched as ONE cluster; every thread takes part in every barrier.
template <typename T>
__global__ void __launch_bounds__(kTailThreads, 1) tail_vcycle_kernel(const TailArgs<T>* __restrict__ ap) {
    cg::cluster_group cluster = cg::this_cluster();
    const TailArgs<T>& A = *ap;
    const int nthreads = (int)(gridDim.x * blockDim.x);
    const int gtid = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int nt = A.nt, D = A.degree;
    const T* cur[kTailMaxLevels];

    for (int t = 0; t + 1 < nt; ++t) {
        const TailLevel<T>& L = A.lv[t];
        const TailLevel<T>& C = A.lv[t + 1];
        const T c0 = L.c2[0];
        if (D == 1) {
            tail_phase<T, false, true>(L, nullptr, L.r, L.xa, c0, T(0), T(0), false, gtid, nthreads);
            cur[t] = L.xa;
        } else {
            tail_phase<T, true, true>(L, nullptr, L.xa, nullptr, c0, L.c1[1], L.c2[1], D > 2, gtid, nthreads);
            cluster.sync();
            T* x = L.xa;
            for (int k = 2; k < D; ++k) {
                T* other = (x == L.xa) ? L.xb : L.xa;
                tail_phase<T, true, false>(L, x, other, nullptr, c0, L.c1[k], L.c2[k], k + 1 < D, gtid, nthreads);
                cluster.sync();
                x = other;
            }
            tail_phase<T, false, false>(L, x, L.r, nullptr, c0, T(0), T(0), false, gtid, nthreads);
            cur[t] = x;
        }
        cluster.sync();
        for (int node = gtid; node < C.n; node += nthreads) {
            const int I = node % C.g.Lx, J = node / C.g.Lx;
            double a0, a1;
            restrict_node<T, true>(L.g, C.g, A.tab, L.r, I, J, a0, a1);
            tail_st2<T>(C.b, node, (T)a0, (T)a1);
        }
        cluster.sync();
    }
    {
        const TailLevel<T>& C = A.lv[nt - 1];
        const int nc = A.nc;
        for (int row = gtid; row < nc; row += nthreads) {
            double s0 = 0.0, s1 = 0.0;
            int k = 0;
            for (; k + 1 < nc; k += 2) {  // Ainv is symmetric: column access is coalesced
                s0 += __ldg(A.Ainv + (size_t)k * nc + row) * (double)__ldcg(C.b + k);
                s1 += __ldg(A.Ainv + (size_t)(k + 1) * nc + row) * (double)__ldcg(C.b + k + 1);
            }
            if (k < nc) s0 += __ldg(A.Ainv + (size_t)k * nc + row) * (double)__ldcg(C.b + k);
            __stcg(C.xa + row, (T)(s0 + s1));
        }
        cur[nt - 1] = C.xa;
        cluster.sync();
    }
    for (int t = nt - 1; t-- > 0;) {
        const TailLevel<T>& L = A.lv[t];
        const TailLevel<T>& C = A.lv[t + 1];
        T* x = const_cast<T*>(cur[t]);
        for (int node = gtid; node < L.n; node += nthreads) {
            const int i = node % L.g.Lx, j = node / L.g.Lx;
            if (L.g.fixed(i, j)) continue;
            double a0, a1;
            prolong_node<T, true>(L.g, C.g, A.tab, cur[t + 1], i, j, a0, a1);
            const typename Vec2<T>::type xv = tail_ld2<T>(x, node);
            tail_st2<T>(x, node, (T)((double)xv.x + a0), (T)((double)xv.y + a1));
        }
        cluster.sync();
        for (int k = 0; k < D; ++k) {
            T* other = (x == L.xa) ? L.xb : L.xa;
            tail_phase<T, true, false>(L, x, other, nullptr, T(0), L.c1[k], L.c2[k], k + 1 < D, gtid, nthreads);
            if (t > 0 || k + 1 < D) cluster.sync();
            x = other;
        }
        cur[t] = x;
    }
}

}  // namespace tmx
