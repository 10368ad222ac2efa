// The tail of the multigrid hierarchy (levels of a few thousand lattice nodes and below) as ONE
// thread-block-cluster kernel per V-cycle.
//
// On those levels a launch does microseconds of work: the march-and-shuffle operator kernel is a
// serial latency chain (loads -> a few hundred dependent fp64 flops on one warp per SM
// sub-partition -> epilogue loads) plus a kernel boundary, ~4 us per launch, ~9 launches per
// level and V-cycle.  Here
//   * the same Galerkin operators are assembled once per hierarchy set-up into node stencils
//     (row of A per node: 2x2 blocks towards the structurally coupled neighbours of its node
//     type - 19 for a vertex, 9 for each kind of edge midpoint - out of the 5x5 neighbourhood);
//   * every CTA of the cluster keeps the stencil rows of the nodes it owns in SHARED memory for
//     the whole V-cycle (cluster.sync flushes L1, so cached global loads would be re-fetched from
//     L2 in every phase; measured: that work, not the barrier, bounded the first version);
//   * a row is split over G lanes (G = 2..32, the smaller the level the more lanes per node), so
//     the per-thread chain is a handful of loads and FMAs and a shuffle reduction;
//   * the phases of the V-cycle (Chebyshev steps, residual, restriction, coarsest solve,
//     prolongation) are separated by the hardware cluster barrier (~0.25 us measured) instead of
//     kernel boundaries; lattice vectors stay in global memory and are read through L2
//     (ld.global.cg) after the barrier's release/acquire.
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
constexpr int kTailThreads = 512;
constexpr int kTailMaxDegree = 8;
constexpr int kTailMaxEntries = 19;  // structural couplings of a vertex node
constexpr int kTailMaxU = 10;        // entries per lane at the smallest split G = 2

template <typename T>
struct TailLevel {
    LevelGeom<T> g;  // replicated geometry: j_off = 0, every row owned
    int n;           // lattice nodes Lx * Ly
    // A task is one warp's work: npt = 32 / G nodes of ONE type (parity of i + 2 * parity of j)
    // times G lanes per node; lane = sub * npt + ln handles entries k = u * G + sub of node ln.
    int G, U, npt;
    int gamma;       // cycles of THIS level per visit of its parent (1 = V-cycle; the coarsest level is exact)
    int nxp[2], nyp[2];  // nodes per parity along x and y
    int toff[5];         // task prefix per type
    int ecnt[4];         // structural entries per type
    signed char ent[4][kTailMaxEntries + 1];  // neighbour index nb = 5*(dj+2) + (di+2) of entry k
    // Task q belongs to CTA q % ncta as its local task q / ncta; coefficient (u, c4) of lane
    // `lane` of local task lt sits at  img_base + ((lt * U + u) * 4 + c4) * 32 + lane  of the
    // CTA's image (c4 = 2*c + c': row component c, column component c').
    int img_base, tasks_per_cta;
    const T* dinv;  // [2n] lattice order, like every vector
    T* b;           // right-hand side (first tail level: written by the launching V-cycle)
    T *xa, *xb, *d, *r;
    T c1[kTailMaxDegree], c2[kTailMaxDegree];  // Chebyshev step k: d = c1 d + c2 D^-1 r
};

template <typename T>
struct TailArgs {
    int nt;        // tail levels; the last one is the coarsest (dense inverse)
    int degree;    // Chebyshev-Jacobi steps before and after the coarse correction
    int nc;        // dofs of the coarsest level
    int dry;       // timing studies: 1 = barriers only, 2 = empty kernel
    int img_len;   // coefficients (T) per CTA image
    const T* img;  // [ncta][img_len]
    const double* Ainv;
    TransferTable tab;
    TailLevel<T> lv[kTailMaxLevels];
};

// host: structural neighbour list of a node type, from the mesh topology (two nodes couple iff
// they share a triangle)
inline int tail_entries(int type, signed char* list) {
    bool on[25] = {};
    const int pi = type & 1, pj = type >> 1;
    for (int qj = pj; qj <= 2; qj += 2)
        for (int qi = pi; qi <= 2; qi += 2) {  // positions of such a node inside its cells
            for (int tt = 0; tt < 2; ++tt) {
                if (cell_to_tri_local(tt, 3 * qj + qi) < 0) continue;
                for (int k = 0; k < 6; ++k) {
                    const int q = tri_local_to_cell(tt, k);
                    on[5 * (q / 3 - qj + 2) + (q % 3 - qi + 2)] = true;
                }
            }
        }
    int cnt = 0;
    for (int nb = 0; nb < 25; ++nb)
        if (on[nb]) list[cnt++] = (signed char)nb;
    return cnt;
}

// host: task layout of one level for a cluster of ncta CTAs; returns the level's image length
template <typename T>
inline int tail_plan_level(TailLevel<T>& V, int ncta, int img_base) {
    const int Lx = V.g.Lx, Ly = V.g.Ly;
    V.n = Lx * Ly;
    const int capacity = ncta * kTailThreads;
    int G = 32;
    while (G > 2 && (long)V.n * G > capacity) G >>= 1;
    V.G = G;
    V.npt = 32 / G;
    V.U = (kTailMaxEntries + G - 1) / G;
    V.nxp[0] = (Lx + 1) / 2; V.nxp[1] = Lx / 2;
    V.nyp[0] = (Ly + 1) / 2; V.nyp[1] = Ly / 2;
    V.toff[0] = 0;
    for (int p = 0; p < 4; ++p) {
        const int cnt = V.nxp[p & 1] * V.nyp[p >> 1];
        V.toff[p + 1] = V.toff[p] + (cnt + V.npt - 1) / V.npt;
        V.ecnt[p] = tail_entries(p, V.ent[p]);
    }
    V.tasks_per_cta = (V.toff[4] + ncta - 1) / ncta;
    V.img_base = img_base;
    return V.tasks_per_cta * V.U * 4 * 32;
}

// task, node-in-task -> type and lattice position; false past the end of the type
template <typename T>
__device__ __forceinline__ bool tail_task_node(const TailLevel<T>& L, int q, int ln, int& type, int& i, int& j) {
    int p = 0;
    while (p < 3 && q >= L.toff[p + 1]) ++p;
    const int w = (q - L.toff[p]) * L.npt + ln, nx = L.nxp[p & 1];
    type = p;
    if (w >= nx * L.nyp[p >> 1]) return false;
    i = 2 * (w % nx) + (p & 1);
    j = 2 * (w / nx) + (p >> 1);
    return true;
}

// Row (node, c) of the assembled level operator, scattered into the CTA images: each of the
// <= 4 cells around the node applies its element matrix to the unit vector of (node, c); by
// symmetry that column is the row.  Dirichlet rows are never used (the phases write zeros
// there), couplings to Dirichlet nodes are dropped (the operator kernels mask the same entries,
// tm_elast.cuh).  Padding entries / lanes of the image stay zero from its allocation.
template <typename T>
__global__ void tail_assemble_kernel(const TailLevel<T> L, int ncta, int img_len, T* __restrict__ img) {
    const LevelGeom<T>& g = L.g;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= L.toff[4] * L.npt * 2) return;
    const int c = t & 1, ln = (t >> 1) % L.npt, q = (t >> 1) / L.npt;
    int type, i, j;
    if (!tail_task_node<T>(L, q, ln, type, i, j)) return;
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
                for (int k = 0; k < 9; ++k) {
                    const bool me = (k == 3 * qj + qi);
                    X[k][0] = (me && c == 0) ? T(1) : T(0);
                    X[k][1] = (me && c == 1) ? T(1) : T(0);
                    Y[k][0] = Y[k][1] = T(0);
                }
                const size_t cidx = (size_t)cy * g.nx + cx;
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    wA[k] = g.W[k * plane + cidx];
                    wB[k] = g.W[(6 + k) * plane + cidx];
                }
                cell_apply<T>(X, wA, wB, g.mat, Y);
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    const int di = k % 3 - qi, dj = k / 3 - qj;
                    if (g.fixed(i + di, j + dj)) continue;
                    const int nb = 5 * (dj + 2) + (di + 2);
                    acc[nb][0] += Y[k][0];
                    acc[nb][1] += Y[k][1];
                }
            }
    }
    T* dst = img + (size_t)(q % ncta) * img_len + L.img_base;
    const int lt = q / ncta;
    for (int k = 0; k < L.ecnt[type]; ++k) {
        const int nb = L.ent[type][k], u = k / L.G, lane = (k % L.G) * L.npt + ln;
        dst[((lt * L.U + u) * 4 + 2 * c + 0) * 32 + lane] = acc[nb][0];
        dst[((lt * L.U + u) * 4 + 2 * c + 1) * 32 + lane] = acc[nb][1];
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

struct TailWho {
    int cta, ncta, warp, nwarps_cta, lane;
};

// One smoothing / residual phase over the level.  CHEB: r = b - A x; d = c1 d + c2 D^-1 r;
// out = x + d.  !CHEB: out = b - A x (and, ONFLY, x itself is stored to xstore).
// ONFLY: x = c0 D^-1 b formed at every neighbour instead of read (the first Chebyshev step from
// a zero guess is never stored).
template <typename T, bool CHEB, bool ONFLY>
__device__ __forceinline__ void tail_phase(const TailLevel<T>& L, const T* __restrict__ simg, const T* cur, T* out,
                                           T* xstore, T c0, T c1, T c2, bool store_d, const TailWho& me) {
    using V2 = typename Vec2<T>::type;
    const int npt = L.npt, G = L.G, U = L.U, Lx = L.g.Lx, Ly = L.g.Ly;
    const int sub = me.lane / npt, ln = me.lane - sub * npt;
    const V2* dinv = reinterpret_cast<const V2*>(L.dinv);
    for (int lt = me.warp; lt < L.tasks_per_cta; lt += me.nwarps_cta) {
        const int q = lt * me.ncta + me.cta;
        if (q >= L.toff[4]) break;
        int type, i = 0, j = 0;
        const bool valid = tail_task_node<T>(L, q, ln, type, i, j);
        const int node = j * Lx + i;
        const bool fixed = valid && L.g.fixed(i, j);
        const bool live = valid && !fixed;
        const bool lead = valid && sub == 0;  // the lane that owns the node's epilogue
        // epilogue operands first: their latency overlaps the row
        V2 bv, dv, dp, xo;
        bv.x = bv.y = dv.x = dv.y = dp.x = dp.y = xo.x = xo.y = T(0);
        if (lead && !fixed) {
            bv = tail_ld2<T>(L.b, node);
            if (CHEB || ONFLY) dv = __ldg(dinv + node);
            if (CHEB && !ONFLY && c1 != T(0)) dp = tail_ld2<T>(L.d, node);
            if (!ONFLY) xo = tail_ld2<T>(cur, node);
        }
        // the lane's entries: neighbour values (all loads in flight together), then the sums
        const int ecnt = L.ecnt[type];
        const T* sc = simg + (size_t)(lt * U) * 128 + me.lane;
        V2 xv[kTailMaxU];
#pragma unroll
        for (int u = 0; u < kTailMaxU; ++u) {
            xv[u].x = xv[u].y = T(0);
            const int k = u * G + sub;
            if (u < U && live && k < ecnt) {
                const int nb = L.ent[type][k];
                const int ii = i + nb % 5 - 2, jj = j + nb / 5 - 2;
                if (ii >= 0 && ii < Lx && jj >= 0 && jj < Ly) {
                    const int m = jj * Lx + ii;
                    if (ONFLY) {
                        const V2 bn = tail_ld2<T>(L.b, m);
                        const V2 dn = __ldg(dinv + m);
                        xv[u].x = c0 * dn.x * bn.x;
                        xv[u].y = c0 * dn.y * bn.y;
                    } else {
                        xv[u] = tail_ld2<T>(cur, m);
                    }
                }
            }
        }
        T y0 = T(0), y1 = T(0);
#pragma unroll
        for (int u = 0; u < kTailMaxU; ++u)
            if (u < U) {
                y0 += sc[(u * 4 + 0) * 32] * xv[u].x + sc[(u * 4 + 1) * 32] * xv[u].y;
                y1 += sc[(u * 4 + 2) * 32] * xv[u].x + sc[(u * 4 + 3) * 32] * xv[u].y;
            }
        for (int o = npt; o < 32; o <<= 1) {  // fixed-order butterfly over the G lanes of a node
            y0 += __shfl_xor_sync(0xffffffffu, y0, o);
            y1 += __shfl_xor_sync(0xffffffffu, y1, o);
        }
        if (!lead) continue;
        if (fixed) {  // identity rows with zero data: everything stays zero
            tail_st2<T>(out, node, T(0), T(0));
            if (CHEB && store_d) tail_st2<T>(L.d, node, T(0), T(0));
            if (!CHEB && ONFLY) tail_st2<T>(xstore, node, T(0), T(0));
            continue;
        }
        if (ONFLY) {
            xo.x = c0 * dv.x * bv.x;
            xo.y = c0 * dv.y * bv.y;
        }
        const T r0 = bv.x - y0, r1 = bv.y - y1;
        if (!CHEB) {
            tail_st2<T>(out, node, r0, r1);
            if (ONFLY) tail_st2<T>(xstore, node, xo.x, xo.y);
        } else {
            T d0 = c2 * dv.x * r0, d1 = c2 * dv.y * r1;
            if (ONFLY) {  // the direction of the un-stored first step is x itself
                d0 += c1 * xo.x;
                d1 += c1 * xo.y;
            } else if (c1 != T(0)) {
                d0 += c1 * dp.x;
                d1 += c1 * dp.y;
            }
            if (store_d) tail_st2<T>(L.d, node, d0, d1);
            tail_st2<T>(out, node, xo.x + d0, xo.y + d1);
        }
    }
}

// b_C = P^T r: G lanes (the coarse level's split) share the 49 taps of a coarse node
template <typename T>
__device__ __forceinline__ void tail_restrict(const TailLevel<T>& F, const TailLevel<T>& C, const TransferTable& tab,
                                              int gwarp, int nwarps, int lane) {
    using V2 = typename Vec2<T>::type;
    const int G = C.G, npt = C.npt, sub = lane / npt, ln = lane - sub * npt;
    const int fLx = F.g.Lx, fLy = F.g.Ly;
    for (int base = gwarp * npt; base < C.n; base += nwarps * npt) {
        const int node = base + ln;
        const bool valid = node < C.n;
        const int I = valid ? node % C.g.Lx : 0, J = valid ? node / C.g.Lx : 0;
        const bool live = valid && !C.g.fixed(I, J);
        const double* rw = tab.Rw[(I & 1) + 2 * (J & 1)];
        double a0 = 0.0, a1 = 0.0;
        for (int tp0 = 0; tp0 < 49; tp0 += 4 * G) {
            V2 v[4];
            double w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int tp = tp0 + u * G + sub;
                v[u].x = v[u].y = T(0);
                w[u] = 0.0;
                if (live && tp < 49) {
                    const int i = 2 * I + tp % 7 - 3, j = 2 * J + tp / 7 - 3;
                    w[u] = rw[tp];
                    if (w[u] != 0.0 && i >= 0 && i < fLx && j >= 0 && j < fLy)
                        v[u] = tail_ld2<T>(F.r, (size_t)j * fLx + i);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                a0 += w[u] * (double)v[u].x;
                a1 += w[u] * (double)v[u].y;
            }
        }
        for (int o = npt; o < 32; o <<= 1) {
            a0 += __shfl_xor_sync(0xffffffffu, a0, o);
            a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        }
        if (valid && sub == 0) tail_st2<T>(C.b, node, (T)a0, (T)a1);
    }
}

// z_tail = cycle(b_tail): launched as ONE cluster; every thread takes part in every barrier.
// Dynamic shared memory: the argument block, then this CTA's stencil image.
// The cycle is the recursion of Engine::cycle run as a state machine (control flow is uniform over the
// cluster): level t+1 is cycled lv[t+1].gamma times per visit of level t, the first time from a zero
// guess, then from its own last iterate (a W-cycle where gamma = 2).  warm = 1 / 2: the first tail
// level starts from the guess in its xa / xb instead of zero (the engine repeating the tail's first
// level); the result then lands in the buffer the guess came in.
template <typename T>
__global__ void __launch_bounds__(kTailThreads, 1) tail_vcycle_kernel(const TailArgs<T>* __restrict__ ap, int warm) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char tail_smem[];
    constexpr int kArgBytes = (int)((sizeof(TailArgs<T>) + 15) / 16 * 16);
    {
        const unsigned* src = reinterpret_cast<const unsigned*>(ap);
        unsigned* dst = reinterpret_cast<unsigned*>(tail_smem);
        for (int k = threadIdx.x; k < (int)(sizeof(TailArgs<T>) / 4); k += blockDim.x) dst[k] = __ldg(src + k);
        __syncthreads();
    }
    const TailArgs<T>& A = *reinterpret_cast<const TailArgs<T>*>(tail_smem);
    T* simg = reinterpret_cast<T*>(tail_smem + kArgBytes);
    TailWho me;
    me.cta = (int)blockIdx.x;
    me.ncta = (int)gridDim.x;
    me.warp = (int)(threadIdx.x >> 5);
    me.nwarps_cta = (int)(blockDim.x >> 5);
    me.lane = (int)(threadIdx.x & 31);
    // warps interleave over the CTAs of the cluster for the node-ordered phases
    const int nwarps = me.ncta * me.nwarps_cta, gwarp = me.warp * me.ncta + me.cta;
    const int nthreads = nwarps * 32, gtid = gwarp * 32 + me.lane;
    const int nt = A.nt, D = A.degree;
    if (A.dry) {
        if (A.dry == 1)
            for (int k = 0; k < (nt - 1) * (2 * D + 2 + (D > 1 ? 1 : 0)) + 1; ++k) cluster.sync();
        return;
    }
    {  // this CTA's stencil rows -> shared memory (16-byte copies; img_len is a multiple of 128)
        const int4* src = reinterpret_cast<const int4*>(A.img + (size_t)me.cta * A.img_len);
        int4* dst = reinterpret_cast<int4*>(simg);
        const int n16 = (int)((size_t)A.img_len * sizeof(T) / 16);
        for (int k = threadIdx.x; k < n16; k += blockDim.x) dst[k] = __ldg(src + k);
        __syncthreads();
    }
    const T* cur[kTailMaxLevels];
    int done[kTailMaxLevels];  // cycles of level t+1 completed in the current visit of level t
    int t = 0;
    bool descending = true, have_guess = warm != 0;
    if (have_guess) cur[0] = warm == 2 ? A.lv[0].xb : A.lv[0].xa;

    while (true) {
        if (descending && t == nt - 1) {  // coarsest level: x = Ainv b, one warp per row
            const TailLevel<T>& C = A.lv[nt - 1];
            const int nc = A.nc;
            for (int row = gwarp; row < nc; row += nwarps) {
                double s = 0.0;
                for (int k = me.lane; k < nc; k += 32)
                    s += __ldg(A.Ainv + (size_t)row * nc + k) * (double)__ldcg(C.b + k);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (me.lane == 0) __stcg(C.xa + row, (T)s);
            }
            cur[nt - 1] = C.xa;
            cluster.sync();
            descending = false;
            --t;
        } else if (descending) {  // pre-smoothing, residual, restriction; then down
            const TailLevel<T>& L = A.lv[t];
            const TailLevel<T>& C = A.lv[t + 1];
            const T* sl = simg + L.img_base;
            const T c0 = L.c2[0];
            if (have_guess) {
                T* x = const_cast<T*>(cur[t]);
                for (int k = 0; k < D; ++k) {
                    T* other = (x == L.xa) ? L.xb : L.xa;
                    tail_phase<T, true, false>(L, sl, x, other, nullptr, T(0), L.c1[k], L.c2[k], k + 1 < D, me);
                    cluster.sync();
                    x = other;
                }
                tail_phase<T, false, false>(L, sl, x, L.r, nullptr, c0, T(0), T(0), false, me);
                cur[t] = x;
            } else if (D == 1) {
                tail_phase<T, false, true>(L, sl, nullptr, L.r, L.xa, c0, T(0), T(0), false, me);
                cur[t] = L.xa;
            } else {
                tail_phase<T, true, true>(L, sl, nullptr, L.xa, nullptr, c0, L.c1[1], L.c2[1], D > 2, me);
                cluster.sync();
                T* x = L.xa;
                for (int k = 2; k < D; ++k) {
                    T* other = (x == L.xa) ? L.xb : L.xa;
                    tail_phase<T, true, false>(L, sl, x, other, nullptr, c0, L.c1[k], L.c2[k], k + 1 < D, me);
                    cluster.sync();
                    x = other;
                }
                tail_phase<T, false, false>(L, sl, x, L.r, nullptr, c0, T(0), T(0), false, me);
                cur[t] = x;
            }
            cluster.sync();
            tail_restrict<T>(L, C, A.tab, gwarp, nwarps, me.lane);
            cluster.sync();
            done[t] = 0;
            have_guess = false;
            ++t;
        } else {  // a cycle of level t+1 has finished, its iterate is cur[t+1]
            ++done[t];
            if (t + 1 < nt - 1 && done[t] < A.lv[t + 1].gamma) {  // cycle the child again, from its iterate
                have_guess = true;
                descending = true;
                ++t;
                continue;
            }
            const TailLevel<T>& L = A.lv[t];
            const TailLevel<T>& C = A.lv[t + 1];
            const T* sl = simg + L.img_base;
            T* x = const_cast<T*>(cur[t]);
            for (int node = gtid; node < L.n; node += nthreads) {
                const int i = node % L.g.Lx, j = node / L.g.Lx;
                if (L.g.fixed(i, j)) continue;
                double a0, a1;
                prolong_node<T, true>(L.g, C.g, A.tab.Pw, cur[t + 1], i, j, a0, a1);
                const typename Vec2<T>::type xv = tail_ld2<T>(x, node);
                tail_st2<T>(x, node, (T)((double)xv.x + a0), (T)((double)xv.y + a1));
            }
            cluster.sync();
            for (int k = 0; k < D; ++k) {
                T* other = (x == L.xa) ? L.xb : L.xa;
                tail_phase<T, true, false>(L, sl, x, other, nullptr, T(0), L.c1[k], L.c2[k], k + 1 < D, me);
                if (t > 0 || k + 1 < D) cluster.sync();
                x = other;
            }
            cur[t] = x;
            if (t == 0) break;
            --t;
        }
    }
}

}  // namespace tmx
