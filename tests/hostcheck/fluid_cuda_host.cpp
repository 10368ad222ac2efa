// The CUDA DRIVER of the fluid path (tmx::FluidSolver with both preconditioners, tmx::CudaTriMG)
// compiled for the CPU through cuda_host_shim.h: same class, same member functions, same launch
// sequence, "device memory" = malloc.  Checks what the element-level host build
// (fluid_host.cpp) cannot: buffer sizes, offsets, level bookkeeping, option handling.
#define TM_HOST_SHIM 1
#include "../../topomax_b200/csrc/tm_fluid_cuda.cuh"

extern "C" {

// returns MINRES iterations (negative if not converged); up = [u | p]; out3 = {relres, objective, 0}
int hc_driver_solve(int nx, int ny, double W, double H, double q, double rmin, double rmax, double visc,
                    const double* rho, const double* g_boundary, double rtol, int maxit, int preconditioner,
                    double* up, double* sens_rhs, double* out3) {
    try {
        tmx::FluidSolver solver(nx, ny, W, H, visc, rmin, rmax, 0);
        if (preconditioner) solver.set_option(tmx::TM_FLUID_OPT_PRECOND, 1.0);
        solver.set_density(rho, q);
        const tmx::MinresResult r = solver.solve(g_boundary, rtol, maxit, up);
        out3[0] = r.relres;
        out3[1] = solver.objective(up);
        out3[2] = 0.0;
        if (sens_rhs) solver.sens_rhs(rho, up, sens_rhs);
        // a second solve on the same object (buffers reused, hierarchy rebuilt) must give the same count
        const tmx::MinresResult r2 = solver.solve(g_boundary, rtol, maxit, up);
        if (r2.iterations != r.iterations) return -100000 - r2.iterations;
        return r.converged ? r.iterations : -r.iterations;
    } catch (const std::exception& e) {
        out3[2] = 1.0;
        return -1;
    }
}
}
