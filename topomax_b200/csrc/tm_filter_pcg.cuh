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

}  // namespace tmx
