// Persistent Jacobi-PCG for the Helmholtz filter  (eps^2 K1 + M1) x = rhs
// (reference: FEM_src/filter.py:27-41 solved by LUSolver("mumps"), FEM_src/pde_solver.py:130-131).
//
// The P1 system is small next to the state system (1/8 of its dofs) and well conditioned, but a
// launch-per-operation PCG spends its time in launch latency and host round trips (~600
// iterations per mirror-descent step).  Here the whole solve is ONE cooperative launch: every
// block keeps iterating, phases are separated by grid-wide barriers, dot products are reduced
// through per-block partials that every block folds in the same fixed order (deterministic and
// identical on all blocks, so the convergence test is grid-uniform and needs no host).
//
// Two barriers per iteration:
//   phase A: p_new = D^-1 r + beta p_old evaluated on the fly at the 7 stencil vertices
//            (ping-pong p buffers), Ap = A p_new, partial p.Ap
//   phase B: x += alpha p, r -= alpha Ap, partial r.r and r.D^-1 r
#pragma once

#include <cooperative_groups.h>

#include "tm_p1.cuh"

namespace tmx {

namespace cg = cooperative_groups;

constexpr int kFilterThreads = 512;

struct FilterPcgArgs {
    P1Geom g;
    double alpha, beta;   // operator alpha K1 + beta M1
    double rtol;
    int maxit;
    double* partA;        // [1][nblocks]
    double* partB;        // [3][nblocks]
    double* result;       // [0] iterations, [1] relres, [2] converged flag
    double* coef;         // optional [2*maxit]: CG alpha_k, beta_k (Lanczos coefficients)
    int coef_cap;
};

// fold nvals per-block partial arrays (stride nblocks) in a fixed order; result to all threads
template <int NV>
__device__ __forceinline__ void fold_partials(const double* part, int nblocks, double (&out)[NV]) {
    __shared__ double s_w[NV][32];
    __shared__ double s_res[NV];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        double s = 0.0;
        for (int b = tid; b < nblocks; b += blockDim.x) s += __ldcg(&part[(size_t)v * nblocks + b]);
        s = warp_sum(s);
        if (lane == 0) s_w[v][wid] = s;
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            double t = lane < nw ? s_w[v][lane] : 0.0;
            t = warp_sum(t);
            if (lane == 0) s_res[v] = t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int v = 0; v < NV; ++v) out[v] = s_res[v];
    __syncthreads();
}

template <int NV>
__device__ __forceinline__ void store_block_partials(double (&val)[NV], double* part, int nblocks) {
    __shared__ double s_w[NV][32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const double s = warp_sum(val[v]);
        if (lane == 0) s_w[v][wid] = s;
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            double t = lane < nw ? s_w[v][lane] : 0.0;
            t = warp_sum(t);
            if (lane == 0) part[(size_t)v * nblocks + blockIdx.x] = t;
        }
    }
    __syncthreads();
}

// x holds the initial guess on entry and the solution on exit; r, Ap, p0, p1 are scratch (n1 each)
template <typename T>
__global__ void __launch_bounds__(kFilterThreads)
filter_pcg_kernel(const FilterPcgArgs a, const T* __restrict__ rhs, const T* __restrict__ dinv,
                  T* x, T* r, T* Ap, T* p0, T* p1) {
    cg::grid_group grid = cg::this_grid();
    const P1Geom& g = a.g;
    const int nb = gridDim.x;
    const size_t n1 = (size_t)(g.nx + 1) * (g.ny + 1);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t first = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const int W1 = g.nx + 1;

    // ---- r = rhs - A x ; partials r.r, r.D^-1 r, rhs.rhs
    {
        double val[3] = {0.0, 0.0, 0.0};
        for (size_t v = first; v < n1; v += stride) {
            const int iy = (int)(v / W1), ix = (int)(v - (size_t)iy * W1);
            const double ax = p1_row_apply(g, a.alpha, a.beta, ix, iy,
                                           [&](size_t j) { return (double)x[j]; });
            const double b = (double)rhs[v];
            const double ri = b - ax;
            r[v] = (T)ri;
            val[0] += ri * ri;
            val[1] += ri * (double)dinv[v] * ri;
            val[2] += b * b;
        }
        store_block_partials<3>(val, a.partB, nb);
    }
    grid.sync();
    double red[3];
    fold_partials<3>(a.partB, nb, red);
    double rr = red[0], rz = red[1];
    const double bb = red[2];
    const double target = a.rtol * a.rtol * bb;
    int it = 0;
    bool converged = !(bb > 0.0) || rr <= target;
    if (!(bb > 0.0)) {
        for (size_t v = first; v < n1; v += stride) x[v] = T(0);
    }
    double beta_cg = 0.0;
    T* pold = p0;
    T* pnew = p1;
    while (!converged && it < a.maxit) {
        // ---- phase A
        {
            double val[1] = {0.0};
            const bool firstit = (it == 0);
            for (size_t v = first; v < n1; v += stride) {
                const int iy = (int)(v / W1), ix = (int)(v - (size_t)iy * W1);
                auto pn = [&](size_t j) {
                    const double z = (double)dinv[j] * (double)r[j];
                    return firstit ? z : z + beta_cg * (double)pold[j];
                };
                const double pv = pn(v);
                const double apv = p1_row_apply(g, a.alpha, a.beta, ix, iy, pn);
                pnew[v] = (T)pv;
                Ap[v] = (T)apv;
                val[0] += pv * apv;
            }
            store_block_partials<1>(val, a.partA, nb);
        }
        grid.sync();
        double pap[1];
        fold_partials<1>(a.partA, nb, pap);
        const double alpha_cg = pap[0] != 0.0 ? rz / pap[0] : 0.0;
        // ---- phase B
        {
            double val[3] = {0.0, 0.0, 0.0};
            for (size_t v = first; v < n1; v += stride) {
                const double ri = (double)r[v] - alpha_cg * (double)Ap[v];
                x[v] = (T)((double)x[v] + alpha_cg * (double)pnew[v]);
                r[v] = (T)ri;
                val[0] += ri * ri;
                val[1] += ri * (double)dinv[v] * ri;
            }
            store_block_partials<3>(val, a.partB, nb);
        }
        grid.sync();
        fold_partials<3>(a.partB, nb, red);
        rr = red[0];
        beta_cg = rz != 0.0 ? red[1] / rz : 0.0;
        rz = red[1];
        if (a.coef && it < a.coef_cap && blockIdx.x == 0 && threadIdx.x == 0) {
            a.coef[2 * it] = alpha_cg;
            a.coef[2 * it + 1] = beta_cg;
        }
        ++it;
        T* t = pold; pold = pnew; pnew = t;
        converged = rr <= target || !(rr == rr);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.result[0] = (double)it;
        a.result[1] = bb > 0.0 ? sqrt(rr / bb) : 0.0;
        a.result[2] = (rr <= target || !(bb > 0.0)) ? 1.0 : 0.0;
    }
}

}  // namespace tmx

namespace tmx {

// Chebyshev-Jacobi semi-iteration for the same system, also one cooperative launch, but with
// ONE grid barrier per iteration and no dot products on the critical path (the residual norm
// is folded only every `check` iterations).  The operator is constant, so the spectral bounds
// [lmin, lmax] of D^-1 A are estimated once per engine.  x ping-pongs between `x` and `xalt`.
struct FilterChebArgs {
    P1Geom g;
    double alpha, beta;
    double lmin, lmax;
    double rtol;
    int maxit, check;
    double* part;    // [2][nblocks]
    double* result;  // [0] iterations, [1] relres, [2] converged flag
};

template <typename T>
__global__ void __launch_bounds__(kFilterThreads)
filter_cheb_kernel(const FilterChebArgs a, const T* __restrict__ rhs, const T* __restrict__ dinv,
                   T* x, T* xalt, T* d) {
    cg::grid_group grid = cg::this_grid();
    const P1Geom& g = a.g;
    const int nb = gridDim.x;
    const size_t n1 = (size_t)(g.nx + 1) * (g.ny + 1);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t first = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const int W1 = g.nx + 1;
    const double theta = 0.5 * (a.lmax + a.lmin), delta = 0.5 * (a.lmax - a.lmin), sigma = theta / delta;
    double rho = 1.0 / sigma;
    T* cur = x;
    T* nxt = xalt;
    double bb = 0.0, rr = 0.0;
    bool converged = false;
    int it = 0;
    for (; it < a.maxit && !converged; ++it) {
        double c1, c2;
        if (it == 0) {
            c1 = 0.0;
            c2 = 1.0 / theta;
        } else {
            const double rho_new = 1.0 / (2.0 * sigma - rho);
            c1 = rho_new * rho;
            c2 = 2.0 * rho_new / delta;
            rho = rho_new;
        }
        const bool check = (it % a.check) == 0;
        double val[2] = {0.0, 0.0};
        for (size_t v = first; v < n1; v += stride) {
            const int iy = (int)(v / W1), ix = (int)(v - (size_t)iy * W1);
            const double ax = p1_row_apply(g, a.alpha, a.beta, ix, iy, [&](size_t j) { return (double)cur[j]; });
            const double b = (double)rhs[v];
            const double r = b - ax;
            const double dn = (it == 0 ? 0.0 : c1 * (double)d[v]) + c2 * (double)dinv[v] * r;
            d[v] = (T)dn;
            nxt[v] = (T)((double)cur[v] + dn);
            if (check) {
                val[0] += r * r;
                val[1] += b * b;
            }
        }
        if (check) store_block_partials<2>(val, a.part, nb);
        grid.sync();
        if (check) {
            double red[2];
            fold_partials<2>(a.part, nb, red);
            rr = red[0];
            bb = red[1];
            // rr is the residual of the iterate we just left; the new one is at least as good
            converged = !(bb > 0.0) || rr <= a.rtol * a.rtol * bb || !(rr == rr);
        }
        T* t = cur; cur = nxt; nxt = t;
    }
    // the solution must end in `x`
    if (cur != x) {
        for (size_t v = first; v < n1; v += stride) x[v] = cur[v];
    }
    if (!(bb > 0.0)) {
        for (size_t v = first; v < n1; v += stride) x[v] = T(0);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.result[0] = (double)it;
        a.result[1] = bb > 0.0 ? sqrt(rr / bb) : 0.0;
        a.result[2] = (!(bb > 0.0) || rr <= a.rtol * a.rtol * bb) ? 1.0 : 0.0;
    }
}


// ---------------------------------------------------------------------------------------
// The same Chebyshev-Jacobi iteration, temporally blocked: S iterations per grid barrier.
//
// The kernel above spends its ~8 us per iteration in the grid barrier and in L2 round trips,
// not in arithmetic.  Here every CTA owns one rectangular tile of the vertex grid and keeps the
// tile extended by a halo of S vertices in SHARED memory (x ping-pong, d, rhs).  S steps run on
// shrinking regions: step m is exact on the tile extended by S - m, because the stencil reaches
// one vertex and everything it reads was exact after step m - 1.  Only then are the owned
// vertices written back (global ping-pong buffers, other CTAs load them as halo) and the grid
// synchronises: one barrier and one global round trip per S iterations, the redundant halo
// work is ~(1 + 2S/w)(1 + 2S/h).  The iterates are bit-for-bit those of the un-blocked kernel
// evaluated with the per-class stencil; convergence is tested every S iterations.
// ---------------------------------------------------------------------------------------
constexpr int kFilterTbThreads = 512;
constexpr int kFilterTbMaxRows = 14;  // vertices per thread (one column segment)

struct FilterTbArgs {
    P1Geom g;
    double coef[9][7];  // rows of alpha K1 + beta M1 by vertex class (3*cls_y + cls_x: first, interior,
                        // last), towards centre, E, W, N, S, NE, SW; 0 where there is no neighbour
    double lmin, lmax, rtol;
    int maxit, S;
    int tiles_x, tiles_y, tw, th;  // tiles_x * tiles_y CTAs, owned tile tw x th vertices
    int rows_per_thread;           // a thread marches this many rows of one column of the extended tile
    double* part;    // [2][nblocks]
    double* result;  // [0] iterations, [1] relres, [2] converged flag
};

// host: stencil row of a vertex class, by the element loop of p1_row_apply
inline void p1_class_stencil(const P1Geom& g, double alpha, double beta, int cls_x, int cls_y, double out[7]) {
    const int off[7][2] = {{0, 0}, {1, 0}, {-1, 0}, {0, 1}, {0, -1}, {1, 1}, {-1, -1}};
    for (int o = 0; o < 7; ++o) out[o] = 0.0;
    // cells around the vertex, relative: cx in {-1, 0}, cy in {-1, 0}, dropped outside the mesh
    for (int cy = -1; cy <= 0; ++cy)
        for (int cx = -1; cx <= 0; ++cx) {
            if ((cls_x == 0 && cx < 0) || (cls_x == 2 && cx == 0)) continue;
            if ((cls_y == 0 && cy < 0) || (cls_y == 2 && cy == 0)) continue;
            for (int type = 0; type < 2; ++type) {
                const int kl = corner_to_tri_local(type, -cx, -cy);
                if (kl < 0) continue;
                for (int k = 0; k < 3; ++k) {
                    const int ax = (k == 2) ? 1 : ((k == 1 && type == 0) ? 1 : 0);
                    const int ay = (k == 2) ? 1 : ((k == 1 && type == 1) ? 1 : 0);
                    const int dx = cx + ax, dy = cy + ay;
                    for (int o = 0; o < 7; ++o)
                        if (off[o][0] == dx && off[o][1] == dy)
                            out[o] += alpha * g.Ke[type][kl][k] + beta * g.Me[kl][k];
                }
            }
        }
}

// Thread layout: thread (col, seg) owns column `col` of the extended tile and the row segment
// [seg * RS, seg * RS + RS); it marches up its segment with a 3-row register window of x (three
// shared-memory loads per vertex instead of seven) and keeps its d and rhs values in registers
// for the whole block of S steps.  x lives in two shared-memory arrays (ping-pong) padded by one
// ring of zeros and tall enough for every segment, so the march is unconditional and
// branch-free: RS is a template parameter, every thread computes all its RS rows with the
// stencil row of its column class (registers), and only the write of a row outside the exact
// region of the step is predicated off.  The kernel is issue-bound: ~20 instructions per vertex.
// The two mesh-boundary rows (y = 0, y = ny), whose stencil differs, are handled outside the
// march by the threads that own them.  The Chebyshev coefficients of every iteration come from
// a table (no divisions in the loop).
struct TbEdgeRow {
    int r;        // row index inside the thread's segment, -1: the thread does not own that row
    double d, b;  // its direction and right-hand side
};

template <typename T, int RS>
__global__ void __launch_bounds__(kFilterTbThreads, 1)
filter_cheb_tb_kernel(const FilterTbArgs a, const T* __restrict__ rhs, const double* __restrict__ c12, T* x,
                      T* xalt, T* d, T* dalt) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) double tb_smem[];
    const P1Geom& g = a.g;
    const int nb = gridDim.x, S = a.S;
    const int W1 = g.nx + 1, H1 = g.ny + 1;
    const int EW = a.tw + 2 * S, EH = a.th + 2 * S;
    const int nseg = (int)blockDim.x / EW;
    const int PW = EW + 2, PE = PW * (nseg * RS + 2);  // padded; tall enough for every segment
    double* X0 = tb_smem;
    double* X1 = X0 + PE;
    __shared__ double s_coef[9][8];
    const int tid = threadIdx.x;
    const int col = tid % EW, seg = tid / EW;
    const int tile_x = blockIdx.x % a.tiles_x, tile_y = blockIdx.x / a.tiles_x;
    const int ox = tile_x * a.tw, oy = tile_y * a.th;  // owned tile origin
    const int ox1 = min(ox + a.tw, W1), oy1 = min(oy + a.th, H1);
    const int ex0 = ox - S, ey0 = oy - S;  // extended tile origin (may be < 0)
    const int gx = ex0 + col, ly_first = seg * RS, gy_first = ey0 + ly_first;
    const bool worker = seg < nseg;
    // rows of this thread that exist: inside the extended tile and the mesh
    const bool col_in = worker && gx >= 0 && gx < W1;
    const int nrows = col_in ? max(0, min(RS, EH - ly_first)) : 0;
    const bool col_owned = gx >= ox && gx < ox1;
    const int clsx = gx <= 0 ? 0 : (gx >= g.nx ? 2 : 1);
    if (tid < 63) s_coef[tid / 7][tid % 7] = a.coef[tid / 7][tid % 7];
    if (tid < 9) s_coef[tid][7] = 1.0 / a.coef[tid][0];
    for (int e = tid; e < 2 * PE; e += blockDim.x) tb_smem[e] = 0.0;
    double cc[8];  // stencil row of an interior-row vertex of this column class
#pragma unroll
    for (int o = 0; o < 7; ++o) cc[o] = a.coef[3 + clsx][o];
    cc[7] = 1.0 / a.coef[3 + clsx][0];
    TbEdgeRow bot, top;
    bot.r = (0 - gy_first >= 0 && 0 - gy_first < nrows) ? 0 - gy_first : -1;
    top.r = (g.ny - gy_first >= 0 && g.ny - gy_first < nrows) ? g.ny - gy_first : -1;
    bot.d = top.d = bot.b = top.b = 0.0;
    double Rr[RS], Dr[RS];
#pragma unroll
    for (int r = 0; r < RS; ++r) {
        const int gy = gy_first + r;
        const bool in = r < nrows && gy >= 0 && gy < H1;
        Rr[r] = in ? (double)rhs[(size_t)gy * W1 + gx] : 0.0;
        Dr[r] = 0.0;
        if (r == bot.r) bot.b = Rr[r];
        if (r == top.r) top.b = Rr[r];
    }
    __syncthreads();
    T *cur = x, *nxt = xalt, *dcur = d, *dnxt = dalt;
    double bb = 0.0, rr = 0.0;
    bool converged = false;
    int it = 0;
    const int p_first = (ly_first + 1) * PW + col + 1;  // this thread's first vertex, padded arrays
    // owned rows of this thread (for the residual norm and the write-back)
    const int o_lo = col_owned ? max(oy - gy_first, 0) : 0, o_hi = col_owned ? min(oy1 - gy_first, nrows) : 0;
    while (it < a.maxit && !converged) {
        // the extended tile of the current iterate: x to shared memory, d to registers
#pragma unroll
        for (int r = 0; r < RS; ++r) {
            const int gy = gy_first + r;
            if (r < nrows && gy >= 0 && gy < H1) {
                const size_t v = (size_t)gy * W1 + gx;
                X0[p_first + r * PW] = (double)__ldcg(cur + v);
                Dr[r] = it == 0 ? 0.0 : (double)__ldcg(dcur + v);
                if (r == bot.r) bot.d = Dr[r];
                if (r == top.r) top.d = Dr[r];
            }
        }
        __syncthreads();
        double* Xin = X0;
        double* Xout = X1;
        double val[2] = {0.0, 0.0};
        for (int m = 1; m <= S; ++m, ++it) {
            const double c1 = c12[2 * it], c2 = c12[2 * it + 1];
            const int grow = S - m;  // exact region: owned tile extended by `grow`, inside the mesh
            const bool col_on = gx >= max(ox - grow, 0) && gx < min(ox1 + grow, W1);
            const int ry0 = max(oy - grow, 0), ry1 = min(oy1 + grow, H1);
            // interior rows 1 .. ny-1 of the region, as a range of this thread's rows
            const int r_lo = col_on ? max(max(ry0, 1) - gy_first, 0) : 0;
            const int r_hi = col_on ? min(min(ry1, g.ny) - gy_first, nrows) : 0;
            const bool last = m == S;
            if (worker) {
                const double* q = Xin + p_first;  // vertex of row r
                double* w = Xout + p_first;
                const double k2 = c2 * cc[7];
                // window: rows -1 (W, C), 0 (W, C, E), +1 (C, E) around the current vertex
                double xm_l = q[-PW - 1], xm_c = q[-PW];
                double x0_l = q[-1], x0_c = q[0], x0_r = q[1];
                double xp_c = q[PW], xp_r = q[PW + 1];
#pragma unroll
                for (int r = 0; r < RS; ++r) {
                    const double ax = cc[0] * x0_c + cc[1] * x0_r + cc[2] * x0_l + cc[3] * xp_c + cc[4] * xm_c +
                                      cc[5] * xp_r + cc[6] * xm_l;
                    const double res = Rr[r] - ax;
                    const double dn = c1 * Dr[r] + k2 * res;
                    const bool on = r >= r_lo && r < r_hi;
                    if (on) {
                        Dr[r] = dn;
                        w[r * PW] = x0_c + dn;
                    }
                    if (last && on && r >= o_lo && r < o_hi) {
                        val[0] += res * res;
                        val[1] += Rr[r] * Rr[r];
                    }
                    if (r + 1 < RS) {  // march one row up
                        xm_l = x0_l;
                        xm_c = x0_c;
                        x0_l = q[(r + 1) * PW - 1];
                        x0_c = xp_c;
                        x0_r = xp_r;
                        xp_c = q[(r + 2) * PW];
                        xp_r = q[(r + 2) * PW + 1];
                    }
                }
            }
            // mesh-boundary rows y = 0 and y = ny: their own stencil rows, straight from shared memory
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                TbEdgeRow& er = side == 0 ? bot : top;
                const int gy = side == 0 ? 0 : g.ny;
                if (er.r >= 0 && col_on && gy >= ry0 && gy < ry1) {
                    const double* c = s_coef[(side == 0 ? 0 : 6) + clsx];
                    const int p = p_first + er.r * PW;
                    const double xc = Xin[p];
                    const double ax = c[0] * xc + c[1] * Xin[p + 1] + c[2] * Xin[p - 1] + c[3] * Xin[p + PW] +
                                      c[4] * Xin[p - PW] + c[5] * Xin[p + PW + 1] + c[6] * Xin[p - PW - 1];
                    const double res = er.b - ax;
                    const double dn = c1 * er.d + c2 * c[7] * res;
                    er.d = dn;
                    Xout[p] = xc + dn;
                    if (last && er.r >= o_lo && er.r < o_hi) {
                        val[0] += res * res;
                        val[1] += er.b * er.b;
                    }
                }
            }
            __syncthreads();
            double* t = Xin; Xin = Xout; Xout = t;
        }
        // owned vertices of the new iterate -> the other global buffers
#pragma unroll
        for (int r = 0; r < RS; ++r) {
            if (r >= o_lo && r < o_hi) {
                const size_t v = (size_t)(gy_first + r) * W1 + gx;
                __stcg(nxt + v, (T)Xin[p_first + r * PW]);
                __stcg(dnxt + v, (T)(r == bot.r ? bot.d : (r == top.r ? top.d : Dr[r])));
            }
        }
        store_block_partials<2>(val, a.part, nb);
        grid.sync();
        double red[2];
        fold_partials<2>(a.part, nb, red);
        rr = red[0];
        bb = red[1];
        // rr is the residual of the iterate before the last step; the new one is at least as good
        converged = !(bb > 0.0) || rr <= a.rtol * a.rtol * bb || !(rr == rr);
        T* t = cur; cur = nxt; nxt = t;
        t = dcur; dcur = dnxt; dnxt = t;
    }
    // the solution must end in `x`
    if (cur != x || !(bb > 0.0)) {
#pragma unroll
        for (int r = 0; r < RS; ++r) {
            if (r >= o_lo && r < o_hi) {
                const size_t v = (size_t)(gy_first + r) * W1 + gx;
                x[v] = (bb > 0.0) ? cur[v] : T(0);
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.result[0] = (double)it;
        a.result[1] = bb > 0.0 ? sqrt(rr / bb) : 0.0;
        a.result[2] = (!(bb > 0.0) || rr <= a.rtol * a.rtol * bb) ? 1.0 : 0.0;
    }
}

}  // namespace tmx
