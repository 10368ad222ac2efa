// Matrix-free penalised stiffness operator  y = K(xi) x  on the vector-P2 lattice
// (replaces dolfin's assembly of FEM_src/elasisity_problem.py:112-118 in
// FEM_src/pde_solver.py:117-119 and its Dirichlet treatment in :125).
//
// Thread mapping ("march and shuffle"): one thread per CELL COLUMN, marching upwards through
// the cell rows of its strip.  Each cell (two triangles, element matrices never formed: see
// tm_element.cuh) is evaluated exactly once, entirely in registers:
//   * contributions to the cell's right-hand lattice column travel to lane+1 by __shfl_up,
//   * contributions to the cell's top lattice row are carried to the next march step,
// so every lattice node is written once, by its owner, with no atomics and no shared memory,
// and the summation order is fixed (bit-reproducible).  A warp owns 31 cell columns; lane 0
// re-evaluates the last column of the warp to its left (32/31 redundant flops), and a strip
// re-evaluates the cell row below it ((H+1)/H).
//
// Loads: a warp reads 65 consecutive (u_x,u_y) pairs per lattice row (coalesced 16-byte
// vector loads); the density is read once per cell row.  The epilogue fuses what follows the
// matvec in the solvers so the vectors are not re-read:
//   EP_PLAIN  y = A x
//   EP_DOT    y = A x, and the grid-wide  x . A x   (PCG)
//   EP_RESID  y = b - A x
//   EP_CHEB   r = b - A x;  d = c1 d + c2 D^-1 r;  y = x + d   (one Chebyshev-Jacobi step)
//   EP_CHEBDOT  EP_CHEB + the dot product b . y of the owned nodes (r . z of the PCG)
//   EP_RESID0 x = c2 D^-1 b evaluated on the fly at the 9 nodes (the first smoothing step from a
//             zero guess, never materialised before the operator); y = b - A x, x stored to d
//   EP_CHEB0  the first TWO Chebyshev-Jacobi steps from a zero guess in one pass: x1 = c0 D^-1 b on the fly
//             (its direction d1 = x1), then r = b - A x1; d = c1 x1 + c2 D^-1 r; y = x1 + d
//             (saves the separate first-step kernel of the coarse levels: a launch and four vector passes)
// Dirichlet nodes are identity rows/columns (symmetric elimination; same solution as the
// reference's bc.apply because the prescribed value is zero).
#pragma once

#include "tm_common.cuh"
#include "tm_tables.h"

namespace tmx {

enum Epilogue { EP_PLAIN = 0, EP_DOT = 1, EP_RESID = 2, EP_CHEB = 3, EP_RESID0 = 4, EP_CHEBDOT = 5, EP_CHEB0 = 6 };
// the variants whose INPUT is formed on the fly from b and D^-1 (scale: c2 for EP_RESID0, c0 for EP_CHEB0)
__host__ __device__ constexpr bool ep_on_the_fly(int ep) { return ep == EP_RESID0 || ep == EP_CHEB0; }
// EP_CHEBDOT: EP_CHEB that also accumulates  b . y  -- on the last smoothing step of a V-cycle b is
// the PCG residual r and y the preconditioned residual z, so r.z costs no extra pass over r and z
__host__ __device__ constexpr bool ep_is_cheb(int ep) { return ep == EP_CHEB || ep == EP_CHEBDOT; }

template <typename T>
struct Vec2;
template <>
struct Vec2<double> {
    using type = double2;
};
template <>
struct Vec2<float> {
    using type = float2;
};

template <typename T>
struct ApplyArgs {
    const T* x;     // input lattice vector
    T* y;           // output lattice vector (must not alias x)
    const T* b;     // EP_RESID / EP_CHEB
    const T* dinv;  // EP_CHEB: inverse diagonal
    T* d;           // EP_CHEB: Chebyshev direction, updated in place
    T c1, c2;       // EP_CHEB coefficients; EP_RESID0: c2 scales D^-1 b
    T c0;           // EP_CHEB0: scale of the on-the-fly first step x1 = c0 D^-1 b
    int store_d;    // EP_CHEB: 0 = the updated direction is not needed any more (last step)
    ReduceScratch rs;
    double* dot_out;  // EP_DOT, EP_CHEBDOT
    int rows_per_strip;
};

constexpr int kApplyWarps = 4;  // warps (column groups) per block

// Epilogue operands of one owned node, loaded into registers BEFORE the element arithmetic of the
// march step so that their HBM latency hides behind it (an ncu source view of the first version
// showed 44 % of all stall samples on the instructions consuming b and D^-1 right after loading
// them; prefetch.global.L1 did not remove that).
template <typename T>
struct EpiOps {
    typename Vec2<T>::type b, dinv, d;
};

template <typename T, int EP>
__device__ __forceinline__ void load_epi_ops(const ApplyArgs<T>& a, size_t n, EpiOps<T>& o) {
    using V2 = typename Vec2<T>::type;
    if (EP == EP_RESID || ep_on_the_fly(EP) || ep_is_cheb(EP)) o.b = reinterpret_cast<const V2*>(a.b)[n];
    if (EP == EP_CHEB0) o.dinv = reinterpret_cast<const V2*>(a.dinv)[n];
    if (ep_is_cheb(EP)) {
        o.dinv = reinterpret_cast<const V2*>(a.dinv)[n];
        if (a.c1 != T(0)) o.d = reinterpret_cast<const V2*>(a.d)[n];
    }
}

// what happens to one owned node once its row of K x is complete (v), x being the input there
template <typename T, int EP>
__device__ __forceinline__ void apply_epilogue(const ApplyArgs<T>& a, size_t n, bool fixed, T v0, T v1,
                                               T x0, T x1, double& dot, const EpiOps<T>& ops) {
    using V2 = typename Vec2<T>::type;
    if (fixed && ep_on_the_fly(EP)) {  // x = c D^-1 b vanishes on Dirichlet nodes (b_D = 0)
        x0 = x1 = v0 = v1 = T(0);
    } else if (fixed) {  // identity row: the registers hold the masked input (0), fetch the raw value
        const V2 xr = reinterpret_cast<const V2*>(a.x)[n];
        x0 = v0 = xr.x;
        x1 = v1 = xr.y;
    }
    V2 out;
    if (EP == EP_PLAIN || EP == EP_DOT) {
        out.x = v0;
        out.y = v1;
        if (EP == EP_DOT) dot += (double)x0 * (double)v0 + (double)x1 * (double)v1;
    } else {
        const V2 bb = ops.b;
        const T r0 = bb.x - v0, r1 = bb.y - v1;
        if (EP == EP_RESID || EP == EP_RESID0) {
            out.x = r0;
            out.y = r1;
            if (EP == EP_RESID0) {
                V2 xo;
                xo.x = x0;
                xo.y = x1;
                reinterpret_cast<V2*>(a.d)[n] = xo;
            }
        } else {
            const V2 di = ops.dinv;
            V2 dd;
            dd.x = a.c2 * di.x * r0;
            dd.y = a.c2 * di.y * r1;
            if (EP == EP_CHEB0) {  // the previous direction IS the on-the-fly first iterate
                dd.x += a.c1 * x0;
                dd.y += a.c1 * x1;
            } else if (a.c1 != T(0)) {
                dd.x += a.c1 * ops.d.x;
                dd.y += a.c1 * ops.d.y;
            }
            if (a.store_d) reinterpret_cast<V2*>(a.d)[n] = dd;
            out.x = x0 + dd.x;
            out.y = x1 + dd.y;
            if (EP == EP_CHEBDOT) dot += (double)bb.x * (double)out.x + (double)bb.y * (double)out.y;
        }
    }
    reinterpret_cast<V2*>(a.y)[n] = out;
}

__device__ __forceinline__ void prefetch_l1(const void* p) {
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

// MINB: resident blocks per SM the register allocation is held to; PF: software prefetch (into
// L1) of the next march step's input rows and of this step's epilogue operands, which is what
// hides the HBM latency at the low occupancy a 170-register fp64 kernel runs at.
template <typename T, bool STORED_W, int EP, int MINB = 2, bool PF = false>
__global__ void __launch_bounds__(kApplyWarps * 32, MINB)
elast_apply_kernel(const LevelGeom<T> g, const ApplyArgs<T> a) {
    using V2 = typename Vec2<T>::type;
    pdl_prologue();
    const int lane = threadIdx.x & 31;
    const int cgroup = blockIdx.x * kApplyWarps + (threadIdx.x >> 5);
    const int ix = cgroup * 31 + lane - 1;  // cell column of this thread
    const int iy0 = blockIdx.y * a.rows_per_strip;
    const int iy1 = min(g.ny, iy0 + a.rows_per_strip);
    const bool warp_active = (cgroup * 31 - 1) < g.nx + 1 && (cgroup * 31) <= g.nx;
    const bool cell_ok = ix >= 0 && ix < g.nx;
    const bool owner = lane >= 1 && ix <= g.nx;  // writes lattice columns 2ix (and 2ix+1)
    const bool own_c1 = owner && ix < g.nx;

    double dot = 0.0;
    const T cfly = EP == EP_CHEB0 ? a.c0 : a.c2;  // scale of an on-the-fly input

    if (warp_active) {
        const V2* __restrict__ xv = reinterpret_cast<const V2*>(a.x);
        const int Lx = g.Lx;
        const int i0 = 2 * ix;
        bool colok[3], colfix[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            colok[c] = ix >= 0 && (i0 + c) <= 2 * g.nx && (cell_ok || c == 0);
            colfix[c] = (i0 + c) <= g.dl || (i0 + c) >= g.dr;
        }

        T X[9][2];  // input on the cell's 3x3 lattice block, zeroed on Dirichlet nodes
#pragma unroll
        for (int q = 0; q < 9; ++q) X[q][0] = X[q][1] = T(0);
        T xiv[4] = {T(0), T(0), T(0), T(0)};  // density at v0 v1 v2 v3 (level 0)
        T carry[2][2] = {{T(0), T(0)}, {T(0), T(0)}};

        const int iy_start = iy0 > 0 ? iy0 - 1 : 0;
        // lattice row 2*iy_start enters as the "top row of the previous step"
        {
            const size_t row = (size_t)(2 * iy_start) * Lx;
            const bool rowfix = g.row_fixed(2 * iy_start);
#pragma unroll
            for (int c = 0; c < 3; ++c)
                if (colok[c]) {
                    V2 v;
                    if (ep_on_the_fly(EP)) {
                        const V2 bv = reinterpret_cast<const V2*>(a.b)[row + i0 + c];
                        const V2 dv = reinterpret_cast<const V2*>(a.dinv)[row + i0 + c];
                        v.x = cfly * dv.x * bv.x;
                        v.y = cfly * dv.y * bv.y;
                    } else {
                        v = xv[row + i0 + c];
                    }
                    const bool f = rowfix || colfix[c];
                    X[6 + c][0] = f ? T(0) : v.x;
                    X[6 + c][1] = f ? T(0) : v.y;
                }
            if (!STORED_W && cell_ok) {
                xiv[2] = g.xi[(size_t)iy_start * (g.nx + 1) + ix];
                xiv[3] = g.xi[(size_t)iy_start * (g.nx + 1) + ix + 1];
            }
        }

        // software pipeline: the two new lattice rows (and densities) of step iy+1 are loaded
        // into registers while step iy computes; epilogue operands are prefetched into L1 one
        // step ahead.  Xn/xin hold the rows 2iy+1, 2iy+2 / vertex row iy+1 of the coming step.
        T Xn[6][2];
        T Dn[6][2];  // EP_RESID0 only: raw D^-1 next to raw b (in Xn); multiplied when consumed
        T xin[2] = {T(0), T(0)};
#pragma unroll
        for (int q = 0; q < 6; ++q) Xn[q][0] = Xn[q][1] = Dn[q][0] = Dn[q][1] = T(0);
        auto load_next = [&](int iy_next) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int j = 2 * iy_next + 1 + r;
                const size_t row = (size_t)j * Lx;
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    if (colok[c]) {  // raw values: nothing may consume them before the next step
                        if (ep_on_the_fly(EP)) {
                            const V2 bv = reinterpret_cast<const V2*>(a.b)[row + i0 + c];
                            const V2 dv = reinterpret_cast<const V2*>(a.dinv)[row + i0 + c];
                            Xn[3 * r + c][0] = bv.x;
                            Xn[3 * r + c][1] = bv.y;
                            Dn[3 * r + c][0] = dv.x;
                            Dn[3 * r + c][1] = dv.y;
                        } else {
                            const V2 v = xv[row + i0 + c];
                            Xn[3 * r + c][0] = v.x;
                            Xn[3 * r + c][1] = v.y;
                        }
                    }
            }
            if (!STORED_W && cell_ok) {
                xin[0] = g.xi[(size_t)(iy_next + 1) * (g.nx + 1) + ix];
                xin[1] = g.xi[(size_t)(iy_next + 1) * (g.nx + 1) + ix + 1];
            }
        };
        auto prefetch_epilogue = [&](int iy_e) {
            if (!(PF && colok[0] && (EP == EP_RESID || (ep_is_cheb(EP) && EP != EP_CHEB0))) || iy_e < iy0) return;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const size_t n = (size_t)(2 * iy_e + r) * Lx + i0;
                prefetch_l1(reinterpret_cast<const V2*>(a.b) + n);
                if (ep_is_cheb(EP)) {
                    prefetch_l1(reinterpret_cast<const V2*>(a.dinv) + n);
                    if (a.c1 != T(0)) prefetch_l1(reinterpret_cast<const V2*>(a.d) + n);
                    // (EP_RESID0 needs no epilogue prefetch: b was just read for the on-the-fly x)
                }
            }
        };
        load_next(iy_start);
        prefetch_epilogue(iy_start);

        for (int iy = iy_start; iy < iy1; ++iy) {
            const int j0 = 2 * iy;
            // shift: previous top row becomes the bottom row, the preloaded rows come in
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                X[c][0] = X[6 + c][0];
                X[c][1] = X[6 + c][1];
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {  // Dirichlet nodes enter the operator as zeros
                const bool rowfix = g.row_fixed(j0 + 1 + r);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const bool f = rowfix || colfix[c];
                    if (ep_on_the_fly(EP)) {
                        X[3 + 3 * r + c][0] = f ? T(0) : cfly * Dn[3 * r + c][0] * Xn[3 * r + c][0];
                        X[3 + 3 * r + c][1] = f ? T(0) : cfly * Dn[3 * r + c][1] * Xn[3 * r + c][1];
                    } else {
                        X[3 + 3 * r + c][0] = f ? T(0) : Xn[3 * r + c][0];
                        X[3 + 3 * r + c][1] = f ? T(0) : Xn[3 * r + c][1];
                    }
                }
            }
            xiv[0] = xiv[2];
            xiv[1] = xiv[3];
            xiv[2] = xin[0];
            xiv[3] = xin[1];
            if (iy + 1 < iy1) {
                load_next(iy + 1);
                prefetch_epilogue(iy + 1);
            }
            // epilogue operands of this step's four owned nodes: in flight during the arithmetic
            // (fine level only: the stored-moment kernels of the coarse levels are better off with
            // the registers, i.e. a third resident block per SM, and load them where they are used)
            EpiOps<T> eo[2][2];
            if (!STORED_W && (EP == EP_RESID || ep_on_the_fly(EP) || ep_is_cheb(EP))) {
                if (iy >= iy0 && owner) {
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        if (!g.owns_row(j0 + r)) continue;
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            if (c == 1 && !own_c1) continue;
                            load_epi_ops<T, EP>(a, (size_t)(j0 + r) * Lx + i0 + c, eo[r][c]);
                        }
                    }
                }
            }
            T acc[9][2];
#pragma unroll
            for (int q = 0; q < 9; ++q) acc[q][0] = acc[q][1] = T(0);
            if (cell_ok) {
                T wA[6], wB[6];
                if (STORED_W) {
                    const size_t plane = (size_t)g.nx * g.ny;
                    const size_t cidx = (size_t)iy * g.nx + ix;
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        wA[k] = g.W[k * plane + cidx];
                        wB[k] = g.W[(6 + k) * plane + cidx];
                    }
                } else {
                    moments_from_xi<T>(xiv[0], xiv[1], xiv[3], g.simp_min, wA);
                    moments_from_xi<T>(xiv[0], xiv[2], xiv[3], g.simp_min, wB);
                }
                cell_apply<T>(X, wA, wB, g.mat, acc);
            }
            // right lattice column of the cell -> owner of that column (lane + 1)
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    T recv = __shfl_up_sync(0xffffffffu, acc[3 * r + 2][k], 1);
                    if (lane == 0) recv = T(0);
                    acc[3 * r][k] += recv;
                }
            // rows j0 (now complete) and j0+1 of the owned columns
            if (iy >= iy0 && owner) {
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int j = j0 + r;
                    if (!g.owns_row(j)) continue;
                    const bool rowfix = g.row_fixed(j);
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        if (c == 1 && !own_c1) continue;
                        T v0 = acc[3 * r + c][0], v1 = acc[3 * r + c][1];
                        if (r == 0) {
                            v0 += carry[c][0];
                            v1 += carry[c][1];
                        }
                        if (STORED_W) load_epi_ops<T, EP>(a, (size_t)j * Lx + i0 + c, eo[r][c]);
                        apply_epilogue<T, EP>(a, (size_t)j * Lx + i0 + c, rowfix || colfix[c], v0, v1,
                                              X[3 * r + c][0], X[3 * r + c][1], dot, eo[r][c]);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                carry[c][0] = acc[6 + c][0];
                carry[c][1] = acc[6 + c][1];
            }
        }
        // the lattice's top row belongs to the last strip
        if (iy1 == g.ny && owner && g.owns_row(2 * g.ny)) {
            const int j = 2 * g.ny;
            const bool rowfix = g.row_fixed(j);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                if (c == 1 && !own_c1) continue;
                EpiOps<T> et;
                load_epi_ops<T, EP>(a, (size_t)j * Lx + i0 + c, et);
                apply_epilogue<T, EP>(a, (size_t)j * Lx + i0 + c, rowfix || colfix[c], carry[c][0],
                                      carry[c][1], X[6 + c][0], X[6 + c][1], dot, et);
            }
        }
    }
    if (EP == EP_DOT || EP == EP_CHEBDOT) {
        double v[1] = {dot};
        double* const outs[1] = {a.dot_out};
        grid_reduce<1>(v, a.rs, outs);
    }
}

// Jacobi diagonal by node-centric gather: dinv = 1 / diag(K), 1 on Dirichlet nodes.
// diag entry of a triangle-local dof = sum_m Q[type][local][comp][m] * w[m] with Q a constant
// table (filled by the host from tri_apply on unit vectors).
template <typename T, bool STORED_W>
__global__ void elast_diag_kernel(const LevelGeom<T> g, const DiagTable tab, T* __restrict__ dinv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.Lx || j >= g.Ly || !g.owns_row(j)) return;
    const size_t n = (size_t)j * g.Lx + i;
    if (g.fixed(i, j)) {
        dinv[2 * n] = T(1);
        dinv[2 * n + 1] = T(1);
        return;
    }
    double d0 = 0.0, d1 = 0.0;
    const int cx_lo = i > 0 ? (i - 1) >> 1 : 0, cx_hi = min(g.nx - 1, i >> 1);
    const int cy_lo = j > 0 ? (j - 1) >> 1 : 0, cy_hi = min(g.ny - 1, j >> 1);
    for (int cy = cy_lo; cy <= cy_hi; ++cy)
        for (int cx = cx_lo; cx <= cx_hi; ++cx) {
            const int q = 3 * (j - 2 * cy) + (i - 2 * cx);
#pragma unroll
            for (int type = 0; type < 2; ++type) {
                const int kl = cell_to_tri_local(type, q);
                if (kl < 0) continue;
                T w[6];
                if (STORED_W) {
                    const size_t plane = (size_t)g.nx * g.ny;
                    const size_t cidx = (size_t)cy * g.nx + cx;
                    for (int k = 0; k < 6; ++k) w[k] = g.W[(6 * type + k) * plane + cidx];
                } else {
                    const T x0 = g.xi[(size_t)cy * (g.nx + 1) + cx];
                    const T x3 = g.xi[(size_t)(cy + 1) * (g.nx + 1) + cx + 1];
                    const T x1 = type == 0 ? g.xi[(size_t)cy * (g.nx + 1) + cx + 1]
                                           : g.xi[(size_t)(cy + 1) * (g.nx + 1) + cx];
                    moments_from_xi<T>(x0, x1, x3, g.simp_min, w);
                }
                for (int m = 0; m < 6; ++m) {
                    d0 += tab.Q[type][kl][0][m] * (double)w[m];
                    d1 += tab.Q[type][kl][1][m] * (double)w[m];
                }
            }
        }
    dinv[2 * n] = T(1.0 / d0);
    dinv[2 * n + 1] = T(1.0 / d1);
}

}  // namespace tmx
