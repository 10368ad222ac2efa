// The CUDA DRIVER of the fluid path (tmx::FluidSolver with both preconditioners, tmx::CudaTriMG)
// compiled for the CPU through cuda_host_shim.h: same class, same member functions, same launch
// sequence, "device memory" = malloc.  Checks what the element-level host build
// (fluid_host.cpp) cannot: buffer sizes, offsets, level bookkeeping, option handling.
#define TM_HOST_SHIM 1
#include <cstdio>
#include "../../topomax_b200/csrc/tm_fluid_cuda.cuh"

extern "C" {

// two successive solves (densities rho1 then rho2) on one solver object, as in an optimisation run;
// returns the iterations of the SECOND solve (negative if not converged), its solution in up
int hc_driver_sequence(int nx, int ny, double W, double H, double q, double rmin, double rmax, double visc,
                       const double* rho1, const double* rho2, const double* g_boundary, double rtol, int maxit,
                       int preconditioner, int warm_start, double* up, double* out3) {
    try {
        tmx::FluidSolver solver(nx, ny, W, H, visc, rmin, rmax, 0);
        if (preconditioner & 1) solver.set_option(tmx::TM_FLUID_OPT_PRECOND, 1.0);
        if (preconditioner & 2) solver.set_option(tmx::TM_FLUID_OPT_DEVICE_SCALARS, 1.0);
        solver.set_option(tmx::TM_FLUID_OPT_WARM_START, warm_start ? 1.0 : 0.0);
        solver.set_density(rho1, q);
        const tmx::MinresResult r1 = solver.solve(g_boundary, rtol, maxit, up);
        out3[0] = r1.iterations;
        solver.set_density(rho2, q);
        const tmx::MinresResult r2 = solver.solve(g_boundary, rtol, maxit, up);
        out3[1] = r2.relres;
        out3[2] = solver.last_solve_was_warm() ? 1.0 : 0.0;
        return r2.converged ? r2.iterations : -r2.iterations;
    } catch (const std::exception& e) {
        out3[2] = -1.0;
        return -1;
    }
}

// returns MINRES iterations (negative if not converged); up = [u | p]; out3 = {relres, objective, 0}
int hc_driver_solve(int nx, int ny, double W, double H, double q, double rmin, double rmax, double visc,
                    const double* rho, const double* g_boundary, double rtol, int maxit, int preconditioner,
                    double* up, double* sens_rhs, double* out3) {
    try {
        tmx::FluidSolver solver(nx, ny, W, H, visc, rmin, rmax, 0);
        if (preconditioner & 1) solver.set_option(tmx::TM_FLUID_OPT_PRECOND, 1.0);
        if (preconditioner & 4) solver.set_option(tmx::TM_FLUID_OPT_DETERMINISTIC, 1.0);  // bit 2: gather kernels
        if (preconditioner & 8) {  // bit 3: six iterations replayed from a captured graph, residual every 12
            solver.set_option(tmx::TM_FLUID_OPT_GRAPH, 1.0);
            solver.set_option(tmx::TM_FLUID_OPT_CHECK_EVERY, 12.0);
        }
        if (preconditioner & 2) {  // bit 1: MINRES scalars resident on the "device", checked every 7 iterations
            solver.set_option(tmx::TM_FLUID_OPT_DEVICE_SCALARS, 1.0);
            solver.set_option(tmx::TM_FLUID_OPT_CHECK_EVERY, 7.0);
        }
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
        std::fprintf(stderr, "hc_driver_solve: %s\n", e.what());
        out3[2] = 1.0;
        return -1;
    }
}

// A mesh the multigrid cannot coarsen far enough: asking for the multigrid preconditioner
// must fail cleanly, at set_option time, and leave a solver that still works with the diagonal one -- twice
// over, since a caught error followed by a retry used to leave a half-planned hierarchy behind.
// returns the iterations of the diagonal solve (negative: not converged), -99999 if nothing threw
int hc_driver_uncoarsenable(int nx, int ny, double W, double H, double q, double rmin, double rmax, double visc,
                            const double* rho, const double* g_boundary, double rtol, int maxit, double* up,
                            double* out3) {
    try {
        tmx::FluidSolver solver(nx, ny, W, H, visc, rmin, rmax, 0);
        int threw = 0;
        for (int attempt = 0; attempt < 2; ++attempt) {
            try {
                solver.set_option(tmx::TM_FLUID_OPT_PRECOND, 1.0);
                solver.set_density(rho, q);
            } catch (const std::exception&) {
                ++threw;
            }
        }
        solver.set_density(rho, q);
        const tmx::MinresResult r = solver.solve(g_boundary, rtol, maxit, up);
        out3[0] = r.relres;
        out3[1] = solver.objective(up);
        out3[2] = threw;
        if (threw != 2) return -99999;
        return r.converged ? r.iterations : -r.iterations;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "hc_driver_uncoarsenable: %s\n", e.what());
        return -1;
    }
}
}
