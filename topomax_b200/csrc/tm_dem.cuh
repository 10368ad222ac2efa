// Q1 strain-energy evaluator of the reference's deep-energy back-end:
//   ObjectiveCalculator.evaluate            DEM_src/objective_calculator.py:51-143
//   StrainEnergy.calculate_strain_energy    DEM_src/elasisity_problem.py:82-97
//   StrainEnergy.calculate_objective_and_gradient / calculate_energy   :99-129
// One fused pass per call: per cell the four Gauss-point strains, sigma:eps, det J, the SIMP weight,
// the density gradient -r'(rho) e and the objective sum r(rho) e (internal energy = half of it);
// a second gather kernel gives d(internal energy)/du, which the reference gets from autograd.
//
// float32 like the reference's torch tensors, and in ITS operation order (no FMA contraction,
// products rounded before the sums) so that per-cell values agree to the last bit or two; the
// 9-digit shape-derivative constants are the reference's (:44-49), not (1 +- 1/sqrt 3)/4.
// Layouts are the reference's: u[(ix*(ny+1)+iy)*2 + c] (``evaluate`` :126-131), density / cell
// arrays [iy][ix].  The element arithmetic is __host__ __device__ and is checked on the CPU against
// the oracle (tests/hostcheck/dem_host.cpp).
#pragma once

#ifdef __CUDACC__
#include "tm_common.cuh"  // ReduceScratch, grid_reduce
#endif
#include "tm_element.cuh"  // TM_HD

namespace tmx {

#ifdef __CUDA_ARCH__
#define TM_FMUL(a, b) __fmul_rn((a), (b))
#define TM_FADD(a, b) __fadd_rn((a), (b))
#else  // host build: compiled with -ffp-contract=off
#define TM_FMUL(a, b) ((a) * (b))
#define TM_FADD(a, b) ((a) + (b))
#endif

struct DemGeom {
    int nx, ny;
    float ddx[4][4], ddy[4][4];  // [gauss point][node]: shape derivative times J^-1 (rounded to float)
    float detj, lam, two_mu;
    float simp_min, one_minus_min, penalty;
};

inline DemGeom dem_make_geom(int nx, int ny, double width, double height, double lam, double mu,
                             double simp_min, double penalty) {
    const double a = 0.394337567, b = 0.105662433;
    // [node][d/ds | d/dt][gauss point]
    const double sd[4][2][4] = {{{-a, -b, -b, -a}, {-a, -a, -b, -b}},
                                {{-b, -a, -a, -b}, {a, a, b, b}},
                                {{a, b, b, a}, {-b, -b, -a, -a}},
                                {{b, a, a, b}, {b, b, a, a}}};
    DemGeom g;
    g.nx = nx;
    g.ny = ny;
    const double dx = width / nx, dy = height / ny;
    const double jx = 1.0 / (dx / 2), jy = 1.0 / (dy / 2);  // np.linalg.inv of the diagonal Jacobian
    for (int q = 0; q < 4; ++q)
        for (int i = 0; i < 4; ++i) {
            g.ddx[q][i] = (float)(jx * sd[i][0][q]);
            g.ddy[q][i] = (float)(jy * sd[i][1][q]);
        }
    g.detj = (float)((dx / 2) * (dy / 2));
    g.lam = (float)lam;
    g.two_mu = (float)(2 * mu);
    g.simp_min = (float)simp_min;
    g.one_minus_min = (float)(1 - simp_min);
    g.penalty = (float)penalty;
    return g;
}

TM_HD float dem_pow(float rho, float p) {
    if (p == 3.0f) return TM_FMUL(TM_FMUL(rho, rho), rho);
    if (p == 2.0f) return TM_FMUL(rho, rho);
    if (p == 1.0f) return rho;
    if (p == 0.0f) return 1.0f;
    return powf(rho, p);
}
// r(rho) = m + rho^p (1 - m),   r'(rho) = p rho^(p-1) (1 - m)      (src/penalizers.py:36-46)
TM_HD float dem_simp(const DemGeom& g, float rho) {
    return TM_FADD(g.simp_min, TM_FMUL(dem_pow(rho, g.penalty), g.one_minus_min));
}
TM_HD float dem_simp_derivative(const DemGeom& g, float rho) {
    return TM_FMUL(TM_FMUL(g.penalty, dem_pow(rho, g.penalty - 1.0f)), g.one_minus_min);
}

// stress at Gauss point q from the 4 corner values (node order N1=(iy,ix), N2=(iy+1,ix),
// N3=(iy,ix+1), N4=(iy+1,ix+1)); returns sigma:eps, optionally sigma
TM_HD float dem_gauss_point(const DemGeom& g, int q, const float (&ux)[4], const float (&uy)[4], float* sig4) {
    float gxx = 0.0f, gxy = 0.0f, gyx = 0.0f, gyy = 0.0f;  // g[component][direction]
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        gxx = TM_FADD(gxx, TM_FMUL(ux[i], g.ddx[q][i]));
        gxy = TM_FADD(gxy, TM_FMUL(ux[i], g.ddy[q][i]));
        gyx = TM_FADD(gyx, TM_FMUL(uy[i], g.ddx[q][i]));
        gyy = TM_FADD(gyy, TM_FMUL(uy[i], g.ddy[q][i]));
    }
    const float exx = TM_FMUL(0.5f, TM_FADD(gxx, gxx));
    const float exy = TM_FMUL(0.5f, TM_FADD(gxy, gyx));
    const float eyy = TM_FMUL(0.5f, TM_FADD(gyy, gyy));
    const float ld = TM_FMUL(g.lam, TM_FADD(gxx, gyy));
    const float sxx = TM_FADD(ld, TM_FMUL(g.two_mu, exx));
    const float sxy = TM_FMUL(g.two_mu, exy);
    const float syy = TM_FADD(ld, TM_FMUL(g.two_mu, eyy));
    if (sig4) {
        sig4[0] = sxx;
        sig4[1] = sxy;
        sig4[2] = sxy;
        sig4[3] = syy;
    }
    // torch.sum(sigma * eps, [0, 1]): (0,0), (0,1), (1,0), (1,1)
    float s = TM_FMUL(sxx, exx);
    s = TM_FADD(s, TM_FMUL(sxy, exy));
    s = TM_FADD(s, TM_FMUL(sxy, exy));
    s = TM_FADD(s, TM_FMUL(syy, eyy));
    return s;
}

// e = det J * sum_q sigma:eps
TM_HD float dem_cell_energy(const DemGeom& g, const float (&ux)[4], const float (&uy)[4]) {
    float e = dem_gauss_point(g, 0, ux, uy, nullptr);
#pragma unroll
    for (int q = 1; q < 4; ++q) e = TM_FADD(e, dem_gauss_point(g, q, ux, uy, nullptr));
    return TM_FMUL(e, g.detj);
}

TM_HD void dem_load_cell(const DemGeom& g, const float* u, int ix, int iy, float (&ux)[4], float (&uy)[4]) {
    const size_t col = (size_t)(g.ny + 1);
    const size_t n1 = (size_t)ix * col + iy, n3 = n1 + col;
    const size_t idx[4] = {n1, n1 + 1, n3, n3 + 1};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        ux[i] = u[2 * idx[i]];
        uy[i] = u[2 * idx[i] + 1];
    }
}

// d(1/2 sum_cells r e)/d u at node (ix, iy): gather over the <= 4 cells around the node
TM_HD void dem_node_gradient(const DemGeom& g, const float* u, const float* density, int ix, int iy,
                             float& gx, float& gy) {
    gx = 0.0f;
    gy = 0.0f;
    // cell (cx, cy) sees the node as local node i: (ix-cx, iy-cy) -> N1 (0,0), N2 (0,1), N3 (1,0), N4 (1,1)
    for (int ox = 0; ox < 2; ++ox)
        for (int oy = 0; oy < 2; ++oy) {
            const int cx = ix - ox, cy = iy - oy;
            if (cx < 0 || cy < 0 || cx >= g.nx || cy >= g.ny) continue;
            const int i = 2 * ox + oy;
            float ux[4], uy[4], sig[4];
            dem_load_cell(g, u, cx, cy, ux, uy);
            const float w = dem_simp(g, density[(size_t)cy * g.nx + cx]) * g.detj;
            float ax = 0.0f, ay = 0.0f;
            for (int q = 0; q < 4; ++q) {
                dem_gauss_point(g, q, ux, uy, sig);
                ax += sig[0] * g.ddx[q][i] + sig[1] * g.ddy[q][i];
                ay += sig[2] * g.ddx[q][i] + sig[3] * g.ddy[q][i];
            }
            gx += w * ax;
            gy += w * ay;
        }
}

#ifdef __CUDACC__
// one thread per cell, iy fastest (the displacement's fastest index); the objective is reduced in
// double with the library's deterministic two-stage reduction
__global__ void __launch_bounds__(256) dem_cell_kernel(DemGeom g, const float* __restrict__ u,
                                                       const float* __restrict__ density,
                                                       float* __restrict__ cell_energy,
                                                       float* __restrict__ grad_density, ReduceScratch rs,
                                                       double* objective) {
    const size_t cells = (size_t)g.nx * g.ny;
    double val[1] = {0.0};
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < cells; t += (size_t)gridDim.x * blockDim.x) {
        const int ix = (int)(t / g.ny), iy = (int)(t - (size_t)ix * g.ny);
        float ux[4], uy[4];
        dem_load_cell(g, u, ix, iy, ux, uy);
        const float e = dem_cell_energy(g, ux, uy);
        const size_t c = (size_t)iy * g.nx + ix;
        const float rho = density[c];
        if (cell_energy) cell_energy[c] = e;
        if (grad_density) grad_density[c] = TM_FMUL(-dem_simp_derivative(g, rho), e);
        val[0] += (double)TM_FMUL(dem_simp(g, rho), e);
    }
    double* const outs[1] = {objective};
    grid_reduce<1>(val, rs, outs);
}

__global__ void __launch_bounds__(256) dem_grad_u_kernel(DemGeom g, const float* __restrict__ u,
                                                         const float* __restrict__ density,
                                                         float* __restrict__ grad_u) {
    const size_t nodes = (size_t)(g.nx + 1) * (g.ny + 1);
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < nodes; t += (size_t)gridDim.x * blockDim.x) {
        const int ix = (int)(t / (g.ny + 1)), iy = (int)(t - (size_t)ix * (g.ny + 1));
        float gx, gy;
        dem_node_gradient(g, u, density, ix, iy, gx, gy);
        grad_u[2 * t] = gx;
        grad_u[2 * t + 1] = gy;
    }
}
#endif

}  // namespace tmx
