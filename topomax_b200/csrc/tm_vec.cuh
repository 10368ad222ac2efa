// Fused vector kernels of the preconditioned conjugate-gradient solvers (replace
// LUSolver("mumps"), reference: FEM_src/pde_solver.py:130-131).  All scalars (alpha, beta,
// norms) live in a small device array `sc` so that an iteration needs no host round trip:
//   sc[0], sc[1]  r.z ping-pong (iteration k reads sc[k&1], writes sc[(k+1)&1])
//   sc[2]         p.Ap
//   sc[3]         r.r
//   sc[4]         b.b
//   sc[5..]       scratch results (dots, integrals)
#pragma once

#include "tm_common.cuh"

namespace tmx {

// SC_RZV: r . z delivered by the V-cycle's last smoothing step (EP_CHEBDOT); a fixed slot, so that
// the captured V-cycle graph can be replayed whichever of SC_RZ0/1 is current
// SC_NEWTON: {c, done, iterations, status} of the device-resident volume projection (md_newton_update_kernel)
enum { SC_RZ0 = 0, SC_RZ1 = 1, SC_PAP = 2, SC_RR = 3, SC_BB = 4, SC_TMP = 5, SC_RZV = 8, SC_FLOOR = 9, SC_NEWTON = 10, SC_COUNT = 16 };

constexpr int kVecThreads = 256;

#define TM_GRID_STRIDE(i, n)                                                 \
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (n); \
         i += (size_t)gridDim.x * blockDim.x)

// x += alpha p;  r -= alpha Ap;  alpha = sc[rz_old] / sc[SC_PAP];  sc[SC_RR] = r.r
// JACOBI: additionally sc[rz_new] = r . (dinv r)   (z is never stored)
template <typename T, bool JACOBI>
__global__ void pcg_update_kernel(size_t n, double* sc, int rz_old, int rz_new, T* __restrict__ x,
                                  T* __restrict__ r, const T* __restrict__ p,
                                  const T* __restrict__ Ap, const T* __restrict__ dinv,
                                  ReduceScratch rs) {
    const double pap = sc[SC_PAP];
    const double alpha = pap != 0.0 ? sc[rz_old] / pap : 0.0;
    double val[2] = {0.0, 0.0};
    TM_GRID_STRIDE(i, n) {
        const double ri = (double)r[i] - alpha * (double)Ap[i];
        x[i] = (T)((double)x[i] + alpha * (double)p[i]);
        r[i] = (T)ri;
        val[0] += ri * ri;
        if (JACOBI) val[1] += ri * ((double)dinv[i] * ri);
    }
    if (JACOBI) {
        double* const outs[2] = {sc + SC_RR, sc + rz_new};
        grid_reduce<2>(val, rs, outs);
    } else {
        double v1[1] = {val[0]};
        double* const outs[1] = {sc + SC_RR};
        grid_reduce<1>(v1, rs, outs);
    }
}

// p = z + beta p,  beta = sc[rz_new] / sc[rz_old];  JACOBI: z = dinv r on the fly.
// rz_store >= 0: sc[rz_new] is also filed under sc[rz_store] for the next iteration (no thread of
// this kernel reads that slot)
template <typename T, bool JACOBI>
__global__ void pcg_direction_kernel(size_t n, double* sc, int rz_old, int rz_new, int rz_store,
                                     T* __restrict__ p, const T* __restrict__ r,
                                     const T* __restrict__ z_or_dinv) {
    const double den = sc[rz_old];
    const double beta = den != 0.0 ? sc[rz_new] / den : 0.0;
    if (rz_store >= 0 && blockIdx.x == 0 && threadIdx.x == 0) sc[rz_store] = sc[rz_new];
    TM_GRID_STRIDE(i, n) {
        const double z = JACOBI ? (double)z_or_dinv[i] * (double)r[i] : (double)z_or_dinv[i];
        p[i] = (T)(z + beta * (double)p[i]);
    }
}

// first direction: p = z (JACOBI: dinv r), sc[rz] = r.z, sc[SC_RR] = r.r
template <typename T, bool JACOBI>
__global__ void pcg_start_kernel(size_t n, double* sc, int rz, T* __restrict__ p,
                                 const T* __restrict__ r, const T* __restrict__ z_or_dinv,
                                 ReduceScratch rs) {
    double val[2] = {0.0, 0.0};
    TM_GRID_STRIDE(i, n) {
        const double ri = (double)r[i];
        const double z = JACOBI ? (double)z_or_dinv[i] * ri : (double)z_or_dinv[i];
        p[i] = (T)z;
        val[0] += ri * ri;
        val[1] += ri * z;
    }
    double* const outs[2] = {sc + SC_RR, sc + rz};
    grid_reduce<2>(val, rs, outs);
}

// *out = a . b
template <typename T>
__global__ void dot_kernel(size_t n, const T* __restrict__ a, const T* __restrict__ b,
                           ReduceScratch rs, double* out) {
    double val[1] = {0.0};
    TM_GRID_STRIDE(i, n) val[0] += (double)a[i] * (double)b[i];
    double* const outs[1] = {out};
    grid_reduce<1>(val, rs, outs);
}

// y = a x + b y
template <typename T>
__global__ void axpby_kernel(size_t n, double a, const T* __restrict__ x, double b,
                             T* __restrict__ y) {
    TM_GRID_STRIDE(i, n) y[i] = (T)(a * (double)x[i] + (b != 0.0 ? b * (double)y[i] : 0.0));
}

// z = a x + b y   (z may alias x or y)
// w_i = +-(eps/2) u_i with a pseudo-random sign per entry (eps = machine epsilon of T): the perturbation
// rounding u to T may cause; K w then samples the residual no solver can resolve (Engine::state_solve)
template <typename T>
__global__ void half_ulp_kernel(size_t n, size_t index_offset, const T* __restrict__ u, T* __restrict__ w) {
    const double half_eps = sizeof(T) == 8 ? 1.1102230246251565e-16 : 5.9604644775390625e-08;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned long long h = (unsigned long long)(i + index_offset) * 0x9E3779B97F4A7C15ull;
        h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
        w[i] = (T)(((h & 1ull) ? half_eps : -half_eps) * (double)u[i]);
    }
}

template <typename T>
__global__ void waxpby_kernel(size_t n, double a, const T* x, double b, const T* y, T* z) {
    TM_GRID_STRIDE(i, n) z[i] = (T)(a * (double)x[i] + b * (double)y[i]);
}

template <typename T>
__global__ void scale_by_kernel(size_t n, const T* __restrict__ d, const T* __restrict__ x,
                                T* __restrict__ y, double s) {
    TM_GRID_STRIDE(i, n) y[i] = (T)(s * (double)d[i] * (double)x[i]);
}

// lattice vector: zero the Dirichlet entries (b_D = 0, reference: FEM_src/pde_solver.py:125)
template <typename T>
__global__ void mask_fixed_kernel(const LevelGeom<T> g, const T* __restrict__ in,
                                  T* __restrict__ out) {
    const size_t n2 = (size_t)g.Lx * g.Ly;
    TM_GRID_STRIDE(n, n2) {
        const int j = (int)(n / g.Lx), i = (int)(n - (size_t)j * g.Lx);
        const bool f = g.fixed(i, j);
        out[2 * n] = f ? T(0) : in[2 * n];
        out[2 * n + 1] = f ? T(0) : in[2 * n + 1];
    }
}

// v = dinv .* t / sqrt(*nrm2)   (one power-iteration step of D^-1 A, normalised)
template <typename T>
__global__ void normalize_scale_kernel(size_t n, const T* __restrict__ dinv, const T* __restrict__ t,
                                       T* __restrict__ v, const double* nrm2) {
    const double s = 1.0 / sqrt(*nrm2);
    TM_GRID_STRIDE(i, n) v[i] = (T)(s * (double)dinv[i] * (double)t[i]);
}

// first Chebyshev step from a zero guess: d = x = s dinv b
template <typename T>
__global__ void cheb_first_kernel(size_t n, double s, const T* __restrict__ dinv,
                                  const T* __restrict__ b, T* __restrict__ d, T* __restrict__ x) {
    pdl_prologue();
    TM_GRID_STRIDE(i, n) {
        const T v = (T)(s * (double)dinv[i] * (double)b[i]);
        d[i] = v;
        x[i] = v;
    }
}

// deterministic start vector for the filter power iteration: mixes smooth and oscillatory parts
template <typename T>
__global__ void fill_alternating_kernel(size_t n, T* __restrict__ v) {
    TM_GRID_STRIDE(i, n) {
        unsigned int h = (unsigned int)i * 2654435761u;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        v[i] = (T)(0.25 + (double)(h & 0xffffu) / 65536.0);
    }
}

template <typename TI, typename TO>
__global__ void convert_kernel(size_t n, const TI* __restrict__ in, TO* __restrict__ out) {
    TM_GRID_STRIDE(i, n) out[i] = (TO)in[i];
}

}  // namespace tmx
