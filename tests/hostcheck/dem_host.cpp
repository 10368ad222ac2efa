// Host build of topomax_b200/csrc/tm_dem.cuh (the same element arithmetic the kernels run), so the
// Q1 strain-energy evaluator can be checked against the oracle and the reference's vectors without a
// GPU.  Test infrastructure only: built by tests/test_dem_oracle.py with g++ -ffp-contract=off.
#include <cstddef>

#include "../../topomax_b200/csrc/tm_dem.cuh"

extern "C" {

// the loops of dem_cell_kernel / dem_grad_u_kernel, serially
double hc_dem_evaluate(int nx, int ny, double width, double height, double lam, double mu, double simp_min,
                       double penalty, const float* u, const float* density, float* cell_energy,
                       float* grad_density, float* grad_u) {
    const tmx::DemGeom g = tmx::dem_make_geom(nx, ny, width, height, lam, mu, simp_min, penalty);
    double objective = 0.0;
    for (int ix = 0; ix < nx; ++ix)
        for (int iy = 0; iy < ny; ++iy) {
            float ux[4], uy[4];
            tmx::dem_load_cell(g, u, ix, iy, ux, uy);
            const float e = tmx::dem_cell_energy(g, ux, uy);
            const size_t c = (size_t)iy * nx + ix;
            const float rho = density[c];
            if (cell_energy) cell_energy[c] = e;
            if (grad_density) grad_density[c] = -tmx::dem_simp_derivative(g, rho) * e;
            objective += (double)(tmx::dem_simp(g, rho) * e);
        }
    if (grad_u)
        for (int ix = 0; ix <= nx; ++ix)
            for (int iy = 0; iy <= ny; ++iy) {
                const size_t t = (size_t)ix * (ny + 1) + iy;
                tmx::dem_node_gradient(g, u, density, ix, iy, grad_u[2 * t], grad_u[2 * t + 1]);
            }
    return objective;
}
}
