// Geometric multigrid for the penalised stiffness operator (the preconditioner that takes
// LUSolver("mumps")'s place, reference: FEM_src/pde_solver.py:130-131).
//
// Hierarchy: P2 spaces on uniformly coarsened right-diagonal meshes are NESTED (each coarse
// triangle is the union of 4 fine triangles of the same family), so the Galerkin coarse
// operator P^T K P is again "the P2 stiffness operator with coefficient r(xi)", and because
// grad(phi_coarse) is affine on a coarse triangle it is determined EXACTLY by six moments
// int r lambda_a lambda_b per coarse triangle.  Coarse levels therefore run the very same
// matrix-free kernel (tm_elast.cuh, STORED_W=true) on 12 numbers per cell, and the moments
// are restricted level to level by a fixed 3x3 congruence per sub-triangle.
//
// Grids whose cell counts are odd are coarsened with ceil(n/2): the coarse mesh overhangs the
// domain, the overhanging fine triangles contribute zero moments, transfers skip the missing
// fine nodes.  The coarse spaces stay nested subspaces (restrictions to the domain).  Fixed
// sides on the far end are pulled in to the last coarse column whose functions vanish on the
// fine Dirichlet nodes (dr' = 2*floor(dr/4)).
#pragma once

#include "tm_common.cuh"
#include "tm_tables.h"

namespace tmx {

// fine moments of triangle `type` of fine cell (cx, cyg) -- cyg is the GLOBAL fine cell row;
// zero outside the fine mesh (overhanging coarse cells)
template <typename T, bool FINE_STORED>
__device__ __forceinline__ void fine_moments(const LevelGeom<T>& f, int type, int cx, int cyg,
                                             double w[6]) {
    const int cy = cyg - (f.j_off >> 1);  // local fine cell row
    if (cx >= f.nx || cyg >= f.nyg || cy < 0 || cy >= f.ny) {
#pragma unroll
        for (int k = 0; k < 6; ++k) w[k] = 0.0;
        return;
    }
    if (FINE_STORED) {
        const size_t plane = (size_t)f.nx * f.ny, cidx = (size_t)cy * f.nx + cx;
#pragma unroll
        for (int k = 0; k < 6; ++k) w[k] = (double)f.W[(6 * type + k) * plane + cidx];
    } else {
        const double x0 = (double)f.xi[(size_t)cy * (f.nx + 1) + cx];
        const double x3 = (double)f.xi[(size_t)(cy + 1) * (f.nx + 1) + cx + 1];
        const double x1 = type == 0 ? (double)f.xi[(size_t)cy * (f.nx + 1) + cx + 1]
                                    : (double)f.xi[(size_t)(cy + 1) * (f.nx + 1) + cx];
        moments_from_xi<double>(x0, x1, x3, (double)f.simp_min, w);
    }
}

// Level-0 moments for a general SIMP exponent (p != 3; tm_element.cuh, "General SIMP exponent"):
// W0 (SoA, 12 planes of nx*ny local cells) from the P1 density, one thread per cell.  The
// fine-level operator, diagonal and coarsening then run their stored-moment variants on it.
template <typename T>
__global__ void mg_fine_moments_kernel(const LevelGeom<T> g, const PenaltySpec ps, T* __restrict__ W0) {
    const int cx = blockIdx.x * blockDim.x + threadIdx.x;
    const int cy = blockIdx.y * blockDim.y + threadIdx.y;
    if (cx >= g.nx || cy >= g.ny) return;
    const size_t plane = (size_t)g.nx * g.ny, cidx = (size_t)cy * g.nx + cx;
    const size_t v0 = (size_t)cy * (g.nx + 1) + cx;
    const double x0 = (double)g.xi[v0], x1 = (double)g.xi[v0 + 1];
    const double x2 = (double)g.xi[v0 + g.nx + 1], x3 = (double)g.xi[v0 + g.nx + 2];
    double w[6];
    moments_general(x0, x1, x3, (double)g.simp_min, ps, w);
#pragma unroll
    for (int k = 0; k < 6; ++k) W0[k * plane + cidx] = (T)w[k];
    moments_general(x0, x2, x3, (double)g.simp_min, ps, w);
#pragma unroll
    for (int k = 0; k < 6; ++k) W0[(6 + k) * plane + cidx] = (T)w[k];
}

// W_c (SoA, 12 planes of nxc*nyc) from the fine level
template <typename T, bool FINE_STORED>
__global__ void mg_coarsen_moments_kernel(const LevelGeom<T> f, int nxc, int nyc, int c_cell_off,
                                          int own_c0, int own_c1, const CoarsenTable tab,
                                          T* __restrict__ Wc) {
    // (nxc, nyc): local coarse cells; local coarse cell row 0 is global row c_cell_off; only the
    // owned rows [own_c0, own_c1) are produced (halo rows arrive by exchange)
    const int I = blockIdx.x * blockDim.x + threadIdx.x;
    const int J = blockIdx.y * blockDim.y + threadIdx.y;
    if (I >= nxc || J >= nyc || J < own_c0 || J >= own_c1) return;
    const int Jg = J + c_cell_off;
    const size_t plane = (size_t)nxc * nyc, cidx = (size_t)J * nxc + I;
    for (int ctype = 0; ctype < 2; ++ctype) {
        double acc[3][3] = {};
        for (int s4 = 0; s4 < 4; ++s4) {
            const int s = 4 * ctype + s4;
            double w[6];
            fine_moments<T, FINE_STORED>(f, tab.type[s], 2 * I + tab.dx[s], 2 * Jg + tab.dy[s], w);
            // acc += M w M^T
            double tmp[3][3];
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    double v = 0.0;
#pragma unroll
                    for (int c = 0; c < 3; ++c) v += tab.M[s][a][c] * w[moment_slot(c, d)];
                    tmp[a][d] = v;
                }
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = a; b < 3; ++b) {
                    double v = 0.0;
#pragma unroll
                    for (int d = 0; d < 3; ++d) v += tmp[a][d] * tab.M[s][b][d];
                    acc[a][b] += v;
                }
        }
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = a; b < 3; ++b)
                Wc[(6 * ctype + moment_slot(a, b)) * plane + cidx] = (T)(0.25 * acc[a][b]);
    }
}

// weight of coarse lattice node (I,J) at fine lattice node (i,j); 0 if unrelated
__device__ __forceinline__ double transfer_weight(const TransferTable& tab, int nxc, int nyc, int i,
                                                  int j, int I, int J) {
    const int cx = min(i >> 2, nxc - 1), cy = min(j >> 2, nyc - 1);
    const int qx = I - 2 * cx, qy = J - 2 * cy;
    if (qx < 0 || qx > 2 || qy < 0 || qy > 2) return 0.0;
    return tab.Pw[5 * (j - 4 * cy) + (i - 4 * cx)][3 * qy + qx];
}

// CG: read through L2 only (ld.global.cg) - for the cluster tail kernel, whose inputs were
// written by other CTAs of the same launch.
template <typename T>
struct MgVec2;
template <>
struct MgVec2<double> {
    using type = double2;
};
template <>
struct MgVec2<float> {
    using type = float2;
};
template <typename T, bool CG>
__device__ __forceinline__ typename MgVec2<T>::type mg_ld2(const T* p, size_t node) {
    using V2 = typename MgVec2<T>::type;
    const V2* q = reinterpret_cast<const V2*>(p) + node;
    return CG ? __ldcg(q) : *q;
}

// (P^T r)(I, J) on the coarse lattice: zero on coarse Dirichlet nodes.  The 7x7 gather is
// unrolled with predicated loads so that all of them are in flight together.
template <typename T, bool CG>
__device__ __forceinline__ void restrict_node(const LevelGeom<T>& f, const LevelGeom<T>& c,
                                              const TransferTable& tab, const T* __restrict__ r, int I,
                                              int J, double& a0, double& a1) {
    a0 = 0.0;
    a1 = 0.0;
    if (c.fixed(I, J)) return;
    const int Jg = J + c.j_off;  // global coarse lattice row
    const double* rw = tab.Rw[(I & 1) + 2 * (Jg & 1)];
    double e0 = 0.0, e1 = 0.0;
#pragma unroll
    for (int dj = -3; dj <= 3; ++dj) {
        const int jg = 2 * Jg + dj;  // global fine lattice row
        const int j = jg - f.j_off;
        const bool rowok = jg >= 0 && jg <= 2 * f.nyg && j >= 0 && j < f.Ly;
        const int jc = min(max(j, 0), f.Ly - 1);
#pragma unroll
        for (int di = -3; di <= 3; ++di) {
            const int i = 2 * I + di;
            const bool ok = rowok && i >= 0 && i < f.Lx;
            const int ic = min(max(i, 0), f.Lx - 1);
            const double w = ok ? rw[7 * (dj + 3) + (di + 3)] : 0.0;
            const size_t n = (size_t)jc * f.Lx + ic;
            typename MgVec2<T>::type v;
            v.x = v.y = T(0);
            if (w != 0.0) v = mg_ld2<T, CG>(r, n);  // predicated: never touches unrelated memory
            if ((dj + di) & 1) {
                e0 += w * (double)v.x;
                e1 += w * (double)v.y;
            } else {
                a0 += w * (double)v.x;
                a1 += w * (double)v.y;
            }
        }
    }
    a0 += e0;
    a1 += e1;
}

// (P xc)(i, j) on the fine lattice (the caller skips fine Dirichlet nodes).  `Pw` = the 25 x 9 weight
// table: TransferTable::Pw itself (kernel parameter space: fine for a few nodes, but an indexed constant
// load per weight serialises over the warp's distinct indices) or a shared-memory copy of it.
template <typename T, bool CG>
__device__ __forceinline__ void prolong_node(const LevelGeom<T>& f, const LevelGeom<T>& c,
                                             const double (*Pw)[9], const T* __restrict__ xc, int i,
                                             int j, double& a0, double& a1) {
    const int jg = j + f.j_off;  // global fine lattice row
    const int cx = min(i >> 2, c.nx - 1), cy = min(jg >> 2, c.nyg - 1);
    const double* pw = Pw[5 * (jg - 4 * cy) + (i - 4 * cx)];
    a0 = 0.0;
    a1 = 0.0;
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        const double w = pw[q];
        const size_t N = (size_t)(2 * cy + q / 3 - c.j_off) * c.Lx + (2 * cx + q % 3);
        typename MgVec2<T>::type v;
        v.x = v.y = T(0);
        if (w != 0.0) v = mg_ld2<T, CG>(xc, N);
        a0 += w * (double)v.x;
        a1 += w * (double)v.y;
    }
}

// bc = P^T r  (zero on coarse Dirichlet nodes)
template <typename T>
__global__ void mg_restrict_kernel(const LevelGeom<T> f, const LevelGeom<T> c,
                                   const TransferTable tab, const T* __restrict__ r,
                                   T* __restrict__ bc) {
    pdl_prologue();
    const int I = blockIdx.x * blockDim.x + threadIdx.x;
    const int J = blockIdx.y * blockDim.y + threadIdx.y;
    if (I >= c.Lx || J >= c.Ly || !c.owns_row(J)) return;
    double a0, a1;
    restrict_node<T, false>(f, c, tab, r, I, J, a0, a1);
    const size_t N = (size_t)J * c.Lx + I;
    bc[2 * N] = (T)a0;
    bc[2 * N + 1] = (T)a1;
}

// The same restriction through a shared-memory tile.  The gather kernel above issues, per warp of 32
// coarse nodes, ~30 predicated loads whose active lanes sit 64 bytes apart (coarse neighbours are two
// fine nodes apart and alternate between the 25-point vertex stencil and the 9-point edge stencils):
// ~480 L1 wavefronts per warp, which is what bounds it (1.8 TB/s at 2e8 dofs, profiles/r2b).  Here a
// block first copies the fine window of its 64 x 8 coarse tile into shared memory with coalesced loads,
// the columns split into four planes by (column mod 4); then every warp handles 32 coarse nodes of ONE
// stencil class (I = I0 + 2 lane + cx), whose reads of a given stencil entry are 32 consecutive 16-byte
// words of one plane: conflict-free, 9 or 25 shared loads per node.  Same weights, same summation order
// as restrict_node: bit-identical results.
// cp.async (LDGSTS): BYTES from global to shared memory without a register round trip; `valid` false
// writes zeros (source size 0: nothing is read, the pointer only has to be a mapped address)
template <int BYTES>
__device__ __forceinline__ void cp_async_zfill(void* smem, const void* gmem, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int src_bytes = valid ? BYTES : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;" ::"r"(s), "l"(gmem), "n"(BYTES), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

constexpr int kRtTI = 64, kRtTJ = 8, kRtThreads = 256;
constexpr int kRtFW = 2 * kRtTI + 5;    // fine columns of the window: 2 I0 - 3 .. 2 (I0 + TI - 1) + 3
constexpr int kRtFH = 2 * kRtTJ + 5;    // fine rows of the window
constexpr int kRtPlane = 34;            // entries per plane and row: ceil(133 / 4), and = 2 mod 8 (bank spread)

// The non-zero entries of the four 7 x 7 restriction stencils as lists of (shared-memory offset, weight),
// split like restrict_node's two accumulators (even / odd dj + di) and kept in its (dj, di) order, so
// that the sums are bit-identical to the gather kernel's.  Offsets are relative to
// &sm[(2 Jr) * 4 kRtPlane + lane] for the class's column parity cx = class & 1.
struct RestrictLists {
    int n[4][2];
    int off[4][2][28];
    double w[4][2][28];
};

inline RestrictLists make_restrict_lists(const TransferTable& tab) {
    RestrictLists L;
    std::memset(&L, 0, sizeof(L));
    for (int cls = 0; cls < 4; ++cls) {
        const int cx = cls & 1;
        for (int dj = -3; dj <= 3; ++dj)
            for (int di = -3; di <= 3; ++di) {
                const double w = tab.Rw[cls][7 * (dj + 3) + (di + 3)];
                if (w == 0.0) continue;
                const int acc = (dj + di) & 1;
                const int m = 2 * cx + di + 3;  // window column = 4 lane + m
                int& k = L.n[cls][acc];
                L.off[cls][acc][k] = (dj + 3) * (4 * kRtPlane) + (m & 3) * kRtPlane + (m >> 2);
                L.w[cls][acc][k] = w;
                ++k;
            }
    }
    return L;
}

// the lists live in constant memory (warp-uniform reads: constant cache, no load/store unit traffic);
// uploaded by the engine when it builds its levels (the same universal numbers for every mesh)
__constant__ RestrictLists c_restrict_lists;

template <typename T>
__global__ void __launch_bounds__(kRtThreads)
mg_restrict_tiled_kernel(const LevelGeom<T> f, const LevelGeom<T> c, const T* __restrict__ r, T* __restrict__ bc) {
    const RestrictLists* lists = &c_restrict_lists;
    using V2 = typename MgVec2<T>::type;
    __shared__ V2 sm[kRtFH * 4 * kRtPlane];
    pdl_prologue();
    const int I0 = blockIdx.x * kRtTI, J0 = blockIdx.y * kRtTJ;
    const int jg_start = 2 * (J0 + c.j_off) - 3;  // global fine row of window row 0
    const int i_start = 2 * I0 - 3;
    // Asynchronous global -> shared copies (cp.async, zero fill outside the lattice): all ~11 copies of a
    // thread are in flight together and cost no registers.  (The first version loaded through registers,
    // one LDG -> STS pair at a time: 62 % of its stall samples sat on that STS and it ran at 2.6 TB/s,
    // profiles/r2d_ncu_restrict_tiled_v1.txt.)
    for (int e = threadIdx.x; e < kRtFH * kRtFW; e += kRtThreads) {
        const int row = e / kRtFW, cc = e - row * kRtFW;
        const int i = i_start + cc, jg = jg_start + row, j = jg - f.j_off;
        const bool ok = i >= 0 && i < f.Lx && jg >= 0 && jg <= 2 * f.nyg && j >= 0 && j < f.Ly;
        const V2* src = reinterpret_cast<const V2*>(r) + (ok ? (size_t)j * f.Lx + i : (size_t)0);
        cp_async_zfill<sizeof(V2)>(&sm[row * (4 * kRtPlane) + (cc & 3) * kRtPlane + (cc >> 2)], src, ok);
    }
    cp_async_wait_all();
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < (2 * kRtTJ) / (kRtThreads / 32); ++k) {
        const int task = warp + (kRtThreads / 32) * k;
        const int Jr = task >> 1, cx = task & 1;
        const int I = I0 + 2 * lane + cx, J = J0 + Jr;
        if (J >= c.Ly) continue;  // warp-uniform
        const int cls = cx + 2 * ((J + c.j_off) & 1);  // one stencil class per warp
        const V2* base = sm + (2 * Jr) * (4 * kRtPlane) + lane;
        double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int n = lists->n[cls][h];
            const int* off = lists->off[cls][h];
            const double* w = lists->w[cls][h];
            double s0 = 0.0, s1 = 0.0;
            for (int q = 0; q < n; ++q) {
                const V2 v = base[off[q]];
                const double wq = w[q];
                s0 += wq * (double)v.x;
                s1 += wq * (double)v.y;
            }
            acc[h][0] = s0;
            acc[h][1] = s1;
        }
        double a0 = acc[0][0] + acc[1][0], a1 = acc[0][1] + acc[1][1];
        if (I < c.Lx && c.owns_row(J)) {
            if (c.fixed(I, J)) a0 = a1 = 0.0;
            V2 out;
            out.x = (T)a0;
            out.y = (T)a1;
            reinterpret_cast<V2*>(bc)[(size_t)J * c.Lx + I] = out;
        }
    }
}

// x += P xc  (fine Dirichlet nodes untouched).  A thread updates kProlongRows fine nodes of one lattice
// column (their coarse loads and x loads are issued before the first use), and the weight table is
// copied to shared memory first: ncu on the 2e8-dof level (profiles/r2i_ncu_prolong_constant_table.txt)
// showed the kernel waiting on its INDEXED CONSTANT loads of the weights (short scoreboard 38 % + MIO
// 10 % of the stall samples; 3.4 TB/s whatever the rows per thread) -- a warp holds four column classes,
// so every `LDC c[0x0][R + ...]` replays four times.  Same arithmetic per node.
constexpr int kProlongRows = 4;

template <typename T>
__global__ void mg_prolong_add_kernel(const LevelGeom<T> f, const LevelGeom<T> c,
                                      const TransferTable tab, const T* __restrict__ xc,
                                      T* __restrict__ x) {
    using V2 = typename MgVec2<T>::type;
    __shared__ double s_pw[25][9];
    for (int k = threadIdx.x + blockDim.x * threadIdx.y; k < 25 * 9; k += blockDim.x * blockDim.y)
        s_pw[k / 9][k % 9] = tab.Pw[k / 9][k % 9];
    __syncthreads();
    pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int jb = (blockIdx.y * blockDim.y + threadIdx.y) * kProlongRows;
    if (i >= f.Lx) return;
    double a0[kProlongRows], a1[kProlongRows];
    V2 v[kProlongRows];
    bool act[kProlongRows];
#pragma unroll
    for (int r = 0; r < kProlongRows; ++r) {
        const int j = jb + r;
        act[r] = j < f.Ly && f.owns_row(j) && !f.fixed(i, j);
        a0[r] = a1[r] = 0.0;
        v[r].x = v[r].y = T(0);
        if (act[r]) {
            v[r] = reinterpret_cast<const V2*>(x)[(size_t)j * f.Lx + i];
            prolong_node<T, false>(f, c, s_pw, xc, i, j, a0[r], a1[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < kProlongRows; ++r)
        if (act[r]) {
            V2 o;
            o.x = (T)((double)v[r].x + a0[r]);
            o.y = (T)((double)v[r].y + a1[r]);
            reinterpret_cast<V2*>(x)[(size_t)(jb + r) * f.Lx + i] = o;
        }
}

// The same prolongation through shared-memory tiles (the large levels).  ncu on the 2e8-dof level showed the
// node-per-thread kernel above bound by its per-node WEIGHT fetches, not by bytes: a warp holds four column
// classes, so every indexed load of a weight replays (constant bank: ADU / IDC 88 % busy, 3.4 TB/s; from
// shared memory: short-scoreboard stalls, 3.5 TB/s; profiles/r2i_ncu_prolong_constant_table.txt).  Here a
// block stages its 128 x 8 tile of x (four planes by column mod 4) and the coarse window under it (two
// planes by column parity) with cp.async, and every warp then updates 32 fine nodes of ONE class
// (a, b) = (i mod 4, j mod 4): i = i0 + 4 lane + a.  Its <= 5 non-zero weights come from a constant-memory
// list with warp-uniform addresses, its coarse and x reads are 32 consecutive 16-byte shared words, and the
// tile goes back with coalesced stores.  A node on a coarse cell's far edge is addressed as class 0 of the
// NEXT cell (i >> 2 unclipped): the interpolation is continuous, the weights are the same numbers in the
// same order, so the result is bit-identical to prolong_node (test_tiled_prolongation_is_bit_identical...).
constexpr int kPtTI = 128, kPtTJ = 8, kPtThreads = 256;
constexpr int kPtCW = 36;                  // entries per coarse-window plane and row (65 columns: 33 + 32; = 4 mod 8)
constexpr int kPtCRows = 7;                // coarse rows under 8 fine rows, whatever the alignment
constexpr int kPtXW = kPtTI / 4 + 2;       // entries per x-tile plane and row (32, padded: = 2 mod 8)

struct ProlongLists {
    int n[16];
    int off[16][6];     // relative to &cw[(2 (cy - cy_lo)) * 2 kPtCW + lane]
    double w[16][6];
};

inline ProlongLists make_prolong_lists(const TransferTable& tab) {
    ProlongLists L;
    std::memset(&L, 0, sizeof(L));
    for (int b = 0; b < 4; ++b)
        for (int a = 0; a < 4; ++a) {
            const int cls = 4 * b + a;
            for (int q = 0; q < 9; ++q) {
                const double w = tab.Pw[5 * b + a][q];
                if (w == 0.0) continue;
                const int qc = q % 3, qr = q / 3;
                int& k = L.n[cls];
                L.off[cls][k] = qr * (2 * kPtCW) + (qc & 1) * kPtCW + (qc >> 1);
                L.w[cls][k] = w;
                ++k;
            }
        }
    return L;
}

__constant__ ProlongLists c_prolong_lists;

template <typename T>
__global__ void __launch_bounds__(kPtThreads)
mg_prolong_tiled_kernel(const LevelGeom<T> f, const LevelGeom<T> c, const T* __restrict__ xc, T* __restrict__ x) {
    using V2 = typename MgVec2<T>::type;
    __shared__ V2 xs[kPtTJ * 4 * kPtXW];
    __shared__ V2 cw[kPtCRows * 2 * kPtCW];
    pdl_prologue();
    const int i0 = blockIdx.x * kPtTI, j0 = blockIdx.y * kPtTJ;
    const int jg0 = j0 + f.j_off;       // global fine row of the tile's first row
    const int cx0 = i0 >> 2, cy_lo = jg0 >> 2;
    // x tile: xs[row][col & 3][col >> 2]
    for (int e = threadIdx.x; e < kPtTJ * kPtTI; e += kPtThreads) {
        const int row = e / kPtTI, col = e - row * kPtTI;
        const int i = i0 + col, j = j0 + row;
        const bool ok = i < f.Lx && j < f.Ly;
        const V2* src = reinterpret_cast<const V2*>(x) + (ok ? (size_t)j * f.Lx + i : (size_t)0);
        cp_async_zfill<sizeof(V2)>(&xs[row * (4 * kPtXW) + (col & 3) * kPtXW + (col >> 2)], src, ok);
    }
    // coarse window: cw[row][col & 1][col >> 1], coarse columns 2 cx0 .. 2 cx0 + 64, rows 2 cy_lo .. + 6
    for (int e = threadIdx.x; e < kPtCRows * 65; e += kPtThreads) {
        const int row = e / 65, col = e - row * 65;
        const int I = 2 * cx0 + col, J = 2 * cy_lo + row - c.j_off;  // local coarse lattice position
        const bool ok = I < c.Lx && J >= 0 && J < c.Ly;
        const V2* src = reinterpret_cast<const V2*>(xc) + (ok ? (size_t)J * c.Lx + I : (size_t)0);
        cp_async_zfill<sizeof(V2)>(&cw[row * (2 * kPtCW) + (col & 1) * kPtCW + (col >> 1)], src, ok);
    }
    cp_async_wait_all();
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < (4 * kPtTJ) / (kPtThreads / 32); ++k) {
        const int task = warp + (kPtThreads / 32) * k;
        const int jr = task >> 2, a = task & 3;
        const int i = i0 + 4 * lane + a, j = j0 + jr;
        if (j >= f.Ly) continue;  // warp-uniform
        const int jg = j + f.j_off;
        const int cls = 4 * (jg & 3) + a;  // one class per warp
        const V2* base = cw + (2 * ((jg >> 2) - cy_lo)) * (2 * kPtCW) + lane;
        const int n = c_prolong_lists.n[cls];
        double a0 = 0.0, a1 = 0.0;
        for (int q = 0; q < n; ++q) {
            const V2 v = base[c_prolong_lists.off[cls][q]];
            const double w = c_prolong_lists.w[cls][q];
            a0 += w * (double)v.x;
            a1 += w * (double)v.y;
        }
        if (i < f.Lx && f.owns_row(j) && !f.fixed(i, j)) {
            V2* xp = &xs[jr * (4 * kPtXW) + a * kPtXW + lane];
            V2 o = *xp;
            o.x = (T)((double)o.x + a0);
            o.y = (T)((double)o.y + a1);
            *xp = o;
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < kPtTJ * kPtTI; e += kPtThreads) {
        const int row = e / kPtTI, col = e - row * kPtTI;
        const int i = i0 + col, j = j0 + row;
        if (i < f.Lx && j < f.Ly && f.owns_row(j))
            reinterpret_cast<V2*>(x)[(size_t)j * f.Lx + i] = xs[row * (4 * kPtXW) + (col & 3) * kPtXW + (col >> 2)];
    }
}

// ---------------------------------------------------------------------------------------
// coarsest level: dense inverse in one block (n = 2*Lx*Ly <= kCoarseMaxDofs)
// ---------------------------------------------------------------------------------------
constexpr int kCoarseMaxDofs = 2 * 9 * 9;

// Ainv (n x n, row-major, symmetric) <- inverse of the dense K of the coarsest level, in ONE block with the
// matrix in shared memory (n <= kCoarseMaxDofs: 210 KB of dynamic shared memory at most): assembly column by
// column, then in-place Gauss-Jordan elimination without pivoting (the matrix is
// symmetric positive definite), 2 block barriers per pivot, and a final symmetrisation.  (The first version,
// a Cholesky factor and one triangular solve pair per column walking global memory from one thread per
// column, took 0.65 ms per kernel for the 90 dofs of a 4 x 2-cell level, once per state solve.)
constexpr int kCoarseInvThreads = 1024;
template <typename T>
__global__ void __launch_bounds__(kCoarseInvThreads, 1)
mg_coarse_inverse_kernel(const LevelGeom<T> g, double* __restrict__ Ainv) {
    extern __shared__ __align__(16) unsigned char coarse_smem[];
    double* A = reinterpret_cast<double*>(coarse_smem);  // n x n
    const int n = 2 * g.Lx * g.Ly;
    double* prow = A + (size_t)n * n;  // pivot row / column of the current step
    double* pcol = prow + n;
    Material<double> mat;
    mat.A11 = (double)g.mat.A11; mat.A22 = (double)g.mat.A22; mat.A12 = (double)g.mat.A12;
    mat.A33 = (double)g.mat.A33; mat.kappa = (double)g.mat.kappa;
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) A[e] = 0.0;
    __syncthreads();
    // column `col` = K applied to the unit vector of (node, comp), cell by cell (<= 4 cells touch a node);
    // a thread owns whole columns, so the additions need no atomics
    for (int col = threadIdx.x; col < n; col += blockDim.x) {
        const int node = col >> 1, comp = col & 1;
        const int ni = node % g.Lx, nj = node / g.Lx;
        if (g.fixed(ni, nj)) {
            A[(size_t)col * n + col] = 1.0;
            continue;
        }
        const int cx_lo = ni > 0 ? (ni - 1) >> 1 : 0, cx_hi = min(g.nx - 1, ni >> 1);
        const int cy_lo = nj > 0 ? (nj - 1) >> 1 : 0, cy_hi = min(g.ny - 1, nj >> 1);
        for (int cy = cy_lo; cy <= cy_hi; ++cy)
            for (int cx = cx_lo; cx <= cx_hi; ++cx) {
                const int li = ni - 2 * cx, lj = nj - 2 * cy;
                double X[9][2] = {}, acc[9][2] = {}, wA[6], wB[6];
                X[3 * lj + li][comp] = 1.0;
                const size_t plane = (size_t)g.nx * g.ny, cidx = (size_t)cy * g.nx + cx;
                for (int k = 0; k < 6; ++k) {
                    wA[k] = (double)g.W[k * plane + cidx];
                    wB[k] = (double)g.W[(6 + k) * plane + cidx];
                }
                cell_apply<double>(X, wA, wB, mat, acc);
                for (int q = 0; q < 9; ++q) {
                    const int i = 2 * cx + q % 3, j = 2 * cy + q / 3;
                    if (g.fixed(i, j)) continue;
                    const size_t rown = (size_t)j * g.Lx + i;
                    A[(2 * rown) * n + col] += acc[q][0];
                    A[(2 * rown + 1) * n + col] += acc[q][1];
                }
            }
    }
    __syncthreads();
    // in-place inverse: after step k the leading (k+1) x (k+1) part holds the partial inverse
    for (int k = 0; k < n; ++k) {
        const double p = 1.0 / A[(size_t)k * n + k];
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            prow[j] = A[(size_t)k * n + j];
            pcol[j] = A[(size_t)j * n + k];
        }
        __syncthreads();
        for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
            const int i = e / n, j = e - i * n;
            double v;
            if (i == k && j == k) v = p;
            else if (i == k) v = prow[j] * p;
            else if (j == k) v = -pcol[i] * p;
            else v = A[e] - pcol[i] * prow[j] * p;
            A[e] = v;
        }
        __syncthreads();
    }
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
        const int i = e / n, j = e - i * n;
        Ainv[e] = 0.5 * (A[e] + A[(size_t)j * n + i]);
    }
}

// x = Ainv b: one thread per row (the coarsest level has <= 162 dofs)
template <typename T>
__global__ void mg_coarse_apply_inverse_kernel(int n, const double* __restrict__ Ainv,
                                               const T* __restrict__ b, T* __restrict__ x) {
    __shared__ double sb[kCoarseMaxDofs];
    for (int i = threadIdx.x; i < n; i += blockDim.x) sb[i] = (double)b[i];
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double* row = Ainv + (size_t)i * n;
        double acc = 0.0;
        for (int k = 0; k < n; ++k) acc += row[k] * sb[k];
        x[i] = (T)acc;
    }
}

// deterministic pseudo-random start vector for the power iteration (zero on Dirichlet nodes)
template <typename T>
__global__ void mg_seed_vector_kernel(const LevelGeom<T> g, T* __restrict__ v) {
    const size_t n2 = (size_t)g.Lx * g.Ly;
    for (size_t n = blockIdx.x * (size_t)blockDim.x + threadIdx.x; n < n2;
         n += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(n / g.Lx), i = (int)(n - (size_t)j * g.Lx);
        const bool f = g.fixed(i, j);
        for (int c = 0; c < 2; ++c) {
            unsigned int h = (unsigned int)(2 * (n + (size_t)g.j_off * g.Lx) + c) * 2654435761u;
            h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
            v[2 * n + c] = f ? T(0) : (T)(0.5 + (double)(h & 0xffffu) / 65536.0);
        }
    }
}

}  // namespace tmx
