// CUDA back-end of the fluid path: launches around the host/device work items of tm_fluid.cuh and
// the driver object behind the tm_fluid_* entries of the C ABI.  (First correct version: see the
// header of tm_fluid.cuh for what is deliberately simple here.)
#pragma once

#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>

#ifdef TM_HOST_SHIM  // serial CPU build of this driver for the CPU tests (tests/hostcheck/cuda_host_shim.h)
#include "cuda_host_shim.h"
#else
#include <cuda_runtime.h>

#include "tm_common.cuh"
#include "tm_vec.cuh"
#endif
#include "tm_fluid.cuh"
#include "tm_trimg_cuda.cuh"

namespace tmx {

struct FluidAtomicAdd {
    __host__ __device__ __forceinline__ void operator()(double* p, double v) const {
#ifdef __CUDA_ARCH__
        atomicAdd(p, v);
#else
        *p += v;
#endif
    }
};

#define TM_FLUID_TRI_LOOP(tid, ntri) \
    for (size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x; tid < (ntri); tid += (size_t)gridDim.x * blockDim.x)

__global__ void __launch_bounds__(128) fluid_mass_kernel(const FluidTables* __restrict__ T, FluidGeom g,
                                                         const double* __restrict__ rho, double* __restrict__ Me,
                                                         size_t ntri) {
    TM_FLUID_TRI_LOOP(tid, ntri) fluid_body_mass(*T, g, rho, Me, ntri, tid);
}

__global__ void __launch_bounds__(128) fluid_apply_kernel(const FluidTables* __restrict__ T, FluidGeom g,
                                                          const double* __restrict__ Me, size_t ntri,
                                                          const double* __restrict__ x, double* y, int mode) {
    TM_FLUID_TRI_LOOP(tid, ntri) fluid_body_apply(*T, g, Me, ntri, x, y, tid, mode, FluidAtomicAdd{});
}

__global__ void __launch_bounds__(128) fluid_diag_kernel(const FluidTables* __restrict__ T, FluidGeom g,
                                                         const double* __restrict__ Me, size_t ntri, double* diag) {
    TM_FLUID_TRI_LOOP(tid, ntri) fluid_body_diag(*T, g, Me, ntri, diag, tid, FluidAtomicAdd{});
}

__global__ void __launch_bounds__(128) fluid_schur_kernel(const FluidTables* __restrict__ T, FluidGeom g, double* diag,
                                                          size_t ntri) {
    TM_FLUID_TRI_LOOP(tid, ntri) fluid_body_schur(*T, g, diag, tid, FluidAtomicAdd{});
}

// diag = 1 on boundary velocity nodes, 0 elsewhere (before the scatter kernels)
__global__ void fluid_diag_init_kernel(FluidGeom g, double* diag, size_t n) {
    const size_t n2 = (size_t)g.Lx * g.Ly;
    TM_GRID_STRIDE(k, n) {
        double v = 0.0;
        if (k < 2 * n2) {
            const size_t node = k >> 1;
            const int j = (int)(node / g.Lx), i = (int)(node - (size_t)j * g.Lx);
            if (i == 0 || j == 0 || i == g.Lx - 1 || j == g.Ly - 1) v = 1.0;
        }
        diag[k] = v;
    }
}

__global__ void __launch_bounds__(128) fluid_objective_kernel(const FluidTables* __restrict__ T, FluidGeom g,
                                                              const double* __restrict__ Me, size_t ntri,
                                                              const double* __restrict__ u, ReduceScratch rs,
                                                              double* out) {
    double val[1] = {0.0};
    TM_FLUID_TRI_LOOP(tid, ntri) val[0] += fluid_body_objective(*T, g, Me, ntri, u, tid);
    double* const outs[1] = {out};
    grid_reduce<1>(val, rs, outs);
}

__global__ void __launch_bounds__(128) fluid_sens_kernel(const FluidTables* __restrict__ T, FluidGeom g,
                                                         const double* __restrict__ rho, const double* __restrict__ u,
                                                         double* out, size_t ntri) {
    TM_FLUID_TRI_LOOP(tid, ntri) fluid_body_sens(*T, g, rho, u, out, tid, FluidAtomicAdd{});
}

// ---- gather forms (opt-in deterministic mode): one work item per output entry
__global__ void __launch_bounds__(128) fluid_apply_gather_kernel(const FluidTables* __restrict__ T, FluidGeom g,
                                                                 const double* __restrict__ Me, size_t ntri,
                                                                 const double* __restrict__ x, double* __restrict__ y,
                                                                 size_t items, int mode) {
    TM_GRID_STRIDE(k, items) fluid_body_apply_gather(*T, g, Me, ntri, x, y, k, mode);
}
__global__ void fluid_diag_gather_kernel(const FluidTables* __restrict__ T, FluidGeom g, const double* __restrict__ Me,
                                         size_t ntri, double* diag, size_t items, int pass) {
    TM_GRID_STRIDE(k, items) fluid_body_diag_gather(*T, g, Me, ntri, diag, k, pass);
}
__global__ void __launch_bounds__(128) fluid_sens_gather_kernel(const FluidTables* __restrict__ T, FluidGeom g,
                                                                const double* __restrict__ rho,
                                                                const double* __restrict__ u, double* __restrict__ out,
                                                                size_t verts) {
    TM_GRID_STRIDE(v, verts) fluid_body_sens_gather(*T, g, rho, u, out, v);
}
__global__ void fluid_pmass_diag_gather_kernel(const FluidTables* __restrict__ T, FluidGeom g, double* diag,
                                               size_t verts) {
    TM_GRID_STRIDE(v, verts) fluid_body_pmass_diag_gather(*T, g, diag, v);
}

// ---- vector kernels of the MINRES loop (combined vector [u | p])
__global__ void fluid_precond_kernel(size_t n, const double* __restrict__ diag, const double* __restrict__ r,
                                     double* __restrict__ z) {
    TM_GRID_STRIDE(i, n) z[i] = r[i] / diag[i];
}
__global__ void fluid_scale_kernel(size_t n, double* x, double s) { TM_GRID_STRIDE(i, n) x[i] *= s; }
__global__ void fluid_axpy2_kernel(size_t n, double* y, double a, const double* __restrict__ p, double b,
                                   const double* __restrict__ q) {
    TM_GRID_STRIDE(i, n) y[i] += a * p[i] + b * q[i];
}
__global__ void fluid_direction_kernel(size_t n, double* __restrict__ wn, const double* __restrict__ z, double a3,
                                       const double* __restrict__ wo, double a2, const double* __restrict__ w,
                                       double inv_a1, double* __restrict__ x, double step) {
    TM_GRID_STRIDE(i, n) {
        const double v = (z[i] - a3 * wo[i] - a2 * w[i]) * inv_a1;
        wn[i] = v;
        x[i] += step * v;
    }
}
// sum of the continuity part of b (for the projection onto the range)
__global__ void fluid_sum_kernel(size_t n, const double* __restrict__ v, ReduceScratch rs, double* out) {
    double val[1] = {0.0};
    TM_GRID_STRIDE(i, n) val[0] += v[i];
    double* const outs[1] = {out};
    grid_reduce<1>(val, rs, outs);
}
// b = -(lifted - mean on the continuity rows)
__global__ void fluid_rhs_kernel(size_t n, size_t nu, double* b, const double* sum, double inv_n1) {
    const double mean = *sum * inv_n1;
    TM_GRID_STRIDE(i, n) b[i] = -(b[i] - (i >= nu ? mean : 0.0));
}
// up = x + [g | 0]
__global__ void fluid_finish_kernel(size_t n, size_t nu, const double* __restrict__ x, const double* __restrict__ g,
                                    double* __restrict__ up) {
    TM_GRID_STRIDE(i, n) up[i] = x[i] + (i < nu ? g[i] : 0.0);
}

// ---- fine-level local matrices of the two multigrid hierarchies (opt-in preconditioner)
__global__ void __launch_bounds__(128) fluid_vel_local_kernel(const FluidTables* __restrict__ T,
                                                              const double* __restrict__ Me, double* __restrict__ L,
                                                              size_t ntri) {
    TM_FLUID_TRI_LOOP(tid, ntri) {
        const int t = (int)(tid & 1);
        int k = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = i; j < 6; ++j, ++k) L[(size_t)k * ntri + tid] = Me[(size_t)k * ntri + tid] + T->Kref[t][i][j];
    }
}
__global__ void __launch_bounds__(128) fluid_darcy_kernel(const FluidTables* __restrict__ T, FluidGeom g,
                                                          const double* __restrict__ rho, double* __restrict__ L,
                                                          size_t ntri) {
    TM_FLUID_TRI_LOOP(tid, ntri) fluid_body_darcy(*T, g, rho, L, ntri, tid);
}
__global__ void __launch_bounds__(128) fluid_pmass_diag_kernel(const FluidTables* __restrict__ T, FluidGeom g,
                                                               double* diag, size_t ntri) {
    TM_FLUID_TRI_LOOP(tid, ntri) fluid_body_pmass_diag(*T, g, diag, tid, FluidAtomicAdd{});
}
__global__ void fluid_pmass_add_kernel(size_t n, const double* __restrict__ diag, const double* __restrict__ r,
                                       double* __restrict__ z) {
    TM_GRID_STRIDE(i, n) z[i] += r[i] / diag[i];
}

// ---- MINRES with device-resident scalars (opt-in): coefficients come from the array S
__global__ void fluid_scalar_init_kernel(double* S, double ref_norm) {
    if (blockIdx.x == 0 && threadIdx.x == 0) minres_scalar_init(S, ref_norm);
}
__global__ void fluid_scalar_a_kernel(double* S) {
    if (blockIdx.x == 0 && threadIdx.x == 0) minres_scalar_a(S);
}
__global__ void fluid_scalar_b_kernel(double* S) {
    if (blockIdx.x == 0 && threadIdx.x == 0) minres_scalar_b(S);
}
__global__ void fluid_scale_dev_kernel(size_t n, double* x, const double* __restrict__ S, int is) {
    const double s = S[is];
    TM_GRID_STRIDE(i, n) x[i] *= s;
}
__global__ void fluid_axpy2_dev_kernel(size_t n, double* y, const double* __restrict__ S, int ia,
                                       const double* __restrict__ p, int ib, const double* __restrict__ q) {
    const double a = S[ia], b = S[ib];
    TM_GRID_STRIDE(i, n) y[i] += a * p[i] + b * q[i];
}
__global__ void fluid_direction_dev_kernel(size_t n, double* __restrict__ wn, const double* __restrict__ z,
                                           const double* __restrict__ S, int i3, const double* __restrict__ wo, int i2,
                                           const double* __restrict__ w, int i1, double* __restrict__ x, int istep) {
    const double a3 = S[i3], a2 = S[i2], inv_a1 = S[i1], step = S[istep];
    TM_GRID_STRIDE(i, n) {
        const double v = (z[i] - a3 * wo[i] - a2 * w[i]) * inv_a1;
        wn[i] = v;
        x[i] += step * v;
    }
}

enum { TM_FLUID_OPT_PRECOND = 1, TM_FLUID_OPT_FINE_STEPS = 2, TM_FLUID_OPT_COARSE_STEPS = 3,
       TM_FLUID_OPT_WARM_START = 4, TM_FLUID_OPT_DEVICE_SCALARS = 5, TM_FLUID_OPT_CHECK_EVERY = 6,
       TM_FLUID_OPT_DETERMINISTIC = 7, TM_FLUID_OPT_GRAPH = 8 };

class FluidSolver {
   public:
    struct DVec {
        double* p = nullptr;
    };
    using Vec = DVec;

    FluidSolver(int nx, int ny, double width, double height, double viscosity, double rmin, double rmax, int device)
        : device_(device) {
        if (nx < 1 || ny < 1 || !(width > 0) || !(height > 0)) throw std::runtime_error("fluid: bad mesh");
        g_.nx = nx; g_.ny = ny; g_.Lx = 2 * nx + 1; g_.Ly = 2 * ny + 1;
        g_.rmin = rmin; g_.rmax = rmax; g_.q = 0.0; g_.viscosity = viscosity;
        ntri_ = (size_t)2 * nx * ny;
        nu_ = (size_t)2 * g_.Lx * g_.Ly;
        n1_ = (size_t)(nx + 1) * (ny + 1);
        n_ = nu_ + n1_;
        TM_CUDA(cudaSetDevice(device_));
        int sms = 148;
        TM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device_));
        max_blocks_ = sms * 16;
        const FluidTables T = fluid_make_tables(width / nx, height / ny);
        TM_CUDA(cudaMalloc(&d_tab_, sizeof(FluidTables)));
        TM_CUDA(cudaMemcpy(d_tab_, &T, sizeof(FluidTables), cudaMemcpyHostToDevice));
        alloc(Me_, 21 * ntri_);
        alloc(diag_, n_);
        alloc(b_, n_);
        alloc(x_, n_);
        alloc(xg_, n_);
        for (auto& w : work_) alloc(w.p, n_);
        rs_.capacity = max_blocks_;
        alloc(rs_.partials, (size_t)2 * rs_.capacity);
        TM_CUDA(cudaMalloc(&rs_.counter, sizeof(unsigned int)));
        TM_CUDA(cudaMemset(rs_.counter, 0, sizeof(unsigned int)));
        alloc(sc_, 8);
        TM_CUDA(cudaMallocHost(&h_sc_, sizeof(double) * 8));
        TM_CUDA(cudaStreamCreate(&own_stream_));
        stream_ = own_stream_;
    }
    ~FluidSolver() {
        cudaSetDevice(device_);
        cudaStreamSynchronize(stream_);
        for (double* p : owned_) cudaFree(p);
        cudaFree(d_tab_);
        cudaFree(rs_.counter);
        cudaFreeHost(h_sc_);
        if (h_ms_) cudaFreeHost(h_ms_);
        drop_graph();
        if (own_stream_) cudaStreamDestroy(own_stream_);
    }
    FluidSolver(const FluidSolver&) = delete;
    FluidSolver& operator=(const FluidSolver&) = delete;

    int device() const { return device_; }
    void set_stream(cudaStream_t s) { stream_ = s ? s : own_stream_; }
    // TM_FLUID_OPT_PRECOND: 0 = diagonal (default), 1 = multigrid (velocity V-cycle | M_p^-1 + Darcy V-cycle)
    void set_option(int opt, double value) {
        switch (opt) {
            case TM_FLUID_OPT_PRECOND:
                if (value != 0.0 && value != 1.0) throw std::runtime_error("fluid: preconditioner must be 0 or 1");
                if (value == 1.0) {  // can this mesh be coarsened far enough?  (throws; nothing has changed yet)
                    CudaTriMG<6>::level_list(TriLevel{g_.nx, g_.ny, 2, 1});
                    CudaTriMG<3>::level_list(TriLevel{g_.nx, g_.ny, 1, 0});
                }
                precond_mode_ = (int)value;
                have_density_ = false;
                break;
            case TM_FLUID_OPT_FINE_STEPS: mg_prm_.fine_steps = std::max(1, (int)value); break;
            case TM_FLUID_OPT_COARSE_STEPS: mg_prm_.coarse_steps = std::max(1, (int)value); break;
            case TM_FLUID_OPT_WARM_START:
                warm_ = value != 0.0;
                have_prev_ = false;
                break;
            case TM_FLUID_OPT_DEVICE_SCALARS: dev_scalars_ = value != 0.0; break;
            case TM_FLUID_OPT_GRAPH:  // six MINRES iterations replayed from a captured graph (implies option 5)
                use_graph_ = value != 0.0;
                if (use_graph_) dev_scalars_ = true;
                break;
            case TM_FLUID_OPT_DETERMINISTIC:  // gather kernels: no atomics, bit-reproducible
                deterministic_ = value != 0.0;
                mg_vel_.set_deterministic(deterministic_);
                mg_prs_.set_deterministic(deterministic_);
                have_density_ = false;
                break;
            case TM_FLUID_OPT_CHECK_EVERY: check_every_ = std::max(1, (int)value); break;
            default: throw std::runtime_error("fluid: unknown option " + std::to_string(opt));
        }
    }

    // weighted mass matrices and the diagonal preconditioner for a density field and penaliser q
    void set_density(const double* rho, double q) {
        if (!(q > 0.0)) throw std::runtime_error("fluid: the penalisation q must be > 0");
        g_.q = q;
        TM_LAUNCH(fluid_mass_kernel, tri_grid(), 128, stream_)(d_tab_, g_, rho, Me_, ntri_);
        TM_CHECK_LAUNCH();
        if (deterministic_) {
            const size_t items = nu_ / 2 + n1_;
            for (int pass = 0; pass < 2; ++pass) {
                TM_LAUNCH(fluid_diag_gather_kernel, vec_grid(), kVecThreads, stream_)(d_tab_, g_, Me_, ntri_, diag_,
                                                                                     items, pass);
                TM_CHECK_LAUNCH();
            }
        } else {
            TM_LAUNCH(fluid_diag_init_kernel, vec_grid(), kVecThreads, stream_)(g_, diag_, n_);
            TM_CHECK_LAUNCH();
            TM_LAUNCH(fluid_diag_kernel, tri_grid(), 128, stream_)(d_tab_, g_, Me_, ntri_, diag_);
            TM_CHECK_LAUNCH();
            TM_LAUNCH(fluid_schur_kernel, tri_grid(), 128, stream_)(d_tab_, g_, diag_, ntri_);
            TM_CHECK_LAUNCH();
        }
        if (precond_mode_ == 1) {
            // one flag for "both hierarchies planned and the pressure mass diagonal allocated": a failure
            // anywhere in here (plan() leaves its object un-planned when it throws) is retried, cleanly,
            // by the next call instead of running on half-built state
            if (!mg_ready_) {
                mg_vel_.plan(TriLevel{g_.nx, g_.ny, 2, 1}, max_blocks_);
                mg_prs_.plan(TriLevel{g_.nx, g_.ny, 1, 0}, max_blocks_);
                if (!mp_diag_) alloc(mp_diag_, n1_);
                if (deterministic_) {
                    TM_LAUNCH(fluid_pmass_diag_gather_kernel, vec_grid(), kVecThreads, stream_)(d_tab_, g_, mp_diag_, n1_);
                } else {
                    TM_LAUNCH(fluid_pmass_diag_kernel, tri_grid(), 128, stream_)(d_tab_, g_, mp_diag_, ntri_);
                }
                TM_CHECK_LAUNCH();
                mg_ready_ = true;
            }
            TM_LAUNCH(fluid_vel_local_kernel, tri_grid(), 128, stream_)(d_tab_, Me_, mg_vel_.level_matrices(0), ntri_);
            TM_CHECK_LAUNCH();
            TM_LAUNCH(fluid_darcy_kernel, tri_grid(), 128, stream_)(d_tab_, g_, rho, mg_prs_.level_matrices(0), ntri_);
            TM_CHECK_LAUNCH();
            mg_vel_.setup(stream_, rs_, sc_, h_sc_);
            mg_prs_.setup(stream_, rs_, sc_, h_sc_);
        }
        have_density_ = true;
    }

    // up = [u | p]: u = boundary values + MINRES solution of the homogeneous-boundary system
    MinresResult solve(const double* boundary_velocity, double rtol, int maxit, double* up) {
        need_density();
        TM_CUDA(cudaMemsetAsync(xg_, 0, n_ * sizeof(double), stream_));
        TM_CUDA(cudaMemcpyAsync(xg_, boundary_velocity, nu_ * sizeof(double), cudaMemcpyDeviceToDevice, stream_));
        apply_mode(xg_, b_, 1);
        TM_LAUNCH(fluid_sum_kernel, vec_grid(), kVecThreads, stream_)(n1_, b_ + nu_, rs_, sc_);
        TM_CHECK_LAUNCH();
        TM_LAUNCH(fluid_rhs_kernel, vec_grid(), kVecThreads, stream_)(n_, nu_, b_, sc_, 1.0 / (double)n1_);
        TM_CHECK_LAUNCH();
        DVec b{b_}, x{x_};
        // warm start (opt-in): solve for the correction of the previous solution, to a tolerance
        // relative to the ORIGINAL right-hand side; dropped when its residual is not smaller
        double ref = 0.0;
        bool warm = false;
        if (warm_ && have_prev_) {
            DVec zb = work(0), ax = work(1), xp{xprev_};
            precond(b, zb);
            ref = std::sqrt(std::fmax(dot(zb, b), 0.0));
            apply(xp, ax);
            axpy2(b, -1.0, ax, 0.0, ax);  // b <- b - A x_prev
            precond(b, zb);
            const double res0 = std::sqrt(std::fmax(dot(zb, b), 0.0));
            if (res0 < ref) {
                warm = true;
            } else {
                axpy2(b, 1.0, ax, 0.0, ax);  // restore the right-hand side
                ref = 0.0;
            }
        }
        if (dev_scalars_ && !ms_) {
            alloc(ms_, MS_COUNT);
            TM_CUDA(cudaMallocHost(&h_ms_, sizeof(double) * MS_COUNT));
        }
        const MinresResult r = use_graph_     ? fluid_minres_graph(*this, b, x, rtol, maxit, check_every_, ref)
                               : dev_scalars_ ? fluid_minres_dev(*this, b, x, rtol, maxit, check_every_, ref)
                                              : fluid_minres(*this, b, x, rtol, maxit, ref);
        if (warm) {
            DVec xp{xprev_};
            axpy2(x, 1.0, xp, 0.0, xp);  // x <- x_prev + correction
        }
        if (warm_) {
            if (!xprev_) alloc(xprev_, n_);
            TM_CUDA(cudaMemcpyAsync(xprev_, x_, n_ * sizeof(double), cudaMemcpyDeviceToDevice, stream_));
            have_prev_ = true;
        }
        last_warm_ = warm;
        TM_LAUNCH(fluid_finish_kernel, vec_grid(), kVecThreads, stream_)(n_, nu_, x_, xg_, up);
        TM_CHECK_LAUNCH();
        TM_CUDA(cudaStreamSynchronize(stream_));
        return r;
    }

    double objective(const double* u) {
        need_density();
        TM_LAUNCH(fluid_objective_kernel, tri_grid(), 128, stream_)(d_tab_, g_, Me_, ntri_, u, rs_, sc_);
        TM_CHECK_LAUNCH();
        return read_scalar();
    }

    // NB the result is handed to OTHER objects (the elasticity engine's filter projects it, on the engine's
    // own stream: two blocking streams order against the legacy stream only, not against each other), so
    // the C entry tm_fluid_sens_rhs completes the stream before it returns (sync()), like solve() and
    // objective() do; the same holds for tm_fluid_apply.
    void sens_rhs(const double* rho, const double* u, double* out) {
        if (!(g_.q > 0.0)) throw std::runtime_error("fluid: set the density / penalisation first");
        if (deterministic_) {
            TM_LAUNCH(fluid_sens_gather_kernel, vec_grid(), 128, stream_)(d_tab_, g_, rho, u, out, n1_);
            TM_CHECK_LAUNCH();
        } else {
            TM_CUDA(cudaMemsetAsync(out, 0, n1_ * sizeof(double), stream_));
            TM_LAUNCH(fluid_sens_kernel, tri_grid(), 128, stream_)(d_tab_, g_, rho, u, out, ntri_);
            TM_CHECK_LAUNCH();
        }
    }
    // completes everything enqueued so far (the C entries that hand a result array back call it)
    void sync() { TM_CUDA(cudaStreamSynchronize(stream_)); }

    // y = Op x (mode 0) or the lifting of boundary values (mode 1); exposed for the parity tests
    void apply_mode(const double* x, double* y, int mode) {
        need_density();
        if (deterministic_) {
            const size_t items = nu_ / 2 + n1_;
            TM_LAUNCH(fluid_apply_gather_kernel, vec_grid(), 128, stream_)(d_tab_, g_, Me_, ntri_, x, y, items, mode);
            TM_CHECK_LAUNCH();
        } else {
            TM_CUDA(cudaMemsetAsync(y, 0, n_ * sizeof(double), stream_));
            TM_LAUNCH(fluid_apply_kernel, tri_grid(), 128, stream_)(d_tab_, g_, Me_, ntri_, x, y, mode);
            TM_CHECK_LAUNCH();
        }
    }
    const double* diagonal() const { return diag_; }
    bool last_solve_was_warm() const { return last_warm_; }
    size_t n() const { return n_; }

    // ---- back-end interface of fluid_minres
    DVec& work(int i) { return work_[i]; }
    void zero(DVec& a) { TM_CUDA(cudaMemsetAsync(a.p, 0, n_ * sizeof(double), stream_)); }
    void copy(const DVec& a, DVec& b) {
        TM_CUDA(cudaMemcpyAsync(b.p, a.p, n_ * sizeof(double), cudaMemcpyDeviceToDevice, stream_));
    }
    void scale(DVec& a, double s) {
        TM_LAUNCH(fluid_scale_kernel, vec_grid(), kVecThreads, stream_)(n_, a.p, s);
        TM_CHECK_LAUNCH();
    }
    void apply(const DVec& x, DVec& y) { apply_mode(x.p, y.p, 0); }
    void precond(const DVec& r, DVec& z) {
        if (precond_mode_ == 1) {
            mg_vel_.precondition(r.p, z.p, mg_prm_);
            mg_prs_.precondition(r.p + nu_, z.p + nu_, mg_prm_);
            TM_LAUNCH(fluid_pmass_add_kernel, vec_grid(), kVecThreads, stream_)(n1_, mp_diag_, r.p + nu_, z.p + nu_);
            TM_CHECK_LAUNCH();
            return;
        }
        TM_LAUNCH(fluid_precond_kernel, vec_grid(), kVecThreads, stream_)(n_, diag_, r.p, z.p);
        TM_CHECK_LAUNCH();
    }
    double dot(const DVec& a, const DVec& b) {
        TM_LAUNCH(dot_kernel<double>, vec_grid(), kVecThreads, stream_)(n_, a.p, b.p, rs_, sc_);
        TM_CHECK_LAUNCH();
        return read_scalar();
    }
    void axpy2(DVec& y, double a, const DVec& p, double b, const DVec& q) {
        TM_LAUNCH(fluid_axpy2_kernel, vec_grid(), kVecThreads, stream_)(n_, y.p, a, p.p, b, q.p);
        TM_CHECK_LAUNCH();
    }
    void direction(DVec& wn, const DVec& z, double a3, const DVec& wo, double a2, const DVec& w, double inv_a1,
                   DVec& x, double step) {
        TM_LAUNCH(fluid_direction_kernel, vec_grid(), kVecThreads, stream_)(n_, wn.p, z.p, a3, wo.p, a2, w.p, inv_a1,
                                                                       x.p, step);
        TM_CHECK_LAUNCH();
    }

    // ---- additional back-end interface of fluid_minres_dev
    void dot_dev(const DVec& a, const DVec& b, int idx) {
        TM_LAUNCH(dot_kernel<double>, vec_grid(), kVecThreads, stream_)(n_, a.p, b.p, rs_, ms_ + idx);
        TM_CHECK_LAUNCH();
    }
    void scalar_init(double ref_norm) {
        TM_LAUNCH(fluid_scalar_init_kernel, 1, 32, stream_)(ms_, ref_norm);
        TM_CHECK_LAUNCH();
    }
    void scalar_a() {
        TM_LAUNCH(fluid_scalar_a_kernel, 1, 32, stream_)(ms_);
        TM_CHECK_LAUNCH();
    }
    void scalar_b() {
        TM_LAUNCH(fluid_scalar_b_kernel, 1, 32, stream_)(ms_);
        TM_CHECK_LAUNCH();
    }
    void scale_dev(DVec& a, int is) {
        TM_LAUNCH(fluid_scale_dev_kernel, vec_grid(), kVecThreads, stream_)(n_, a.p, ms_, is);
        TM_CHECK_LAUNCH();
    }
    void axpy2_dev(DVec& y, int ia, const DVec& p, int ib, const DVec& q) {
        TM_LAUNCH(fluid_axpy2_dev_kernel, vec_grid(), kVecThreads, stream_)(n_, y.p, ms_, ia, p.p, ib, q.p);
        TM_CHECK_LAUNCH();
    }
    void direction_dev(DVec& wn, const DVec& z, int i3, const DVec& wo, int i2, const DVec& w, int i1, DVec& x,
                       int istep) {
        TM_LAUNCH(fluid_direction_dev_kernel, vec_grid(), kVecThreads, stream_)(n_, wn.p, z.p, ms_, i3, wo.p, i2, w.p,
                                                                               i1, x.p, istep);
        TM_CHECK_LAUNCH();
    }
    // ---- stream capture of the iteration block of fluid_minres_graph
    void begin_capture() { TM_CUDA(cudaStreamBeginCapture(stream_, cudaStreamCaptureModeThreadLocal)); }
    void end_capture() {
        cudaGraph_t graph = nullptr;
        TM_CUDA(cudaStreamEndCapture(stream_, &graph));
        drop_graph();
        const cudaError_t e = cudaGraphInstantiate(&graph_exec_, graph, 0);
        cudaGraphDestroy(graph);
        TM_CUDA(e);
    }
    void replay() { TM_CUDA(cudaGraphLaunch(graph_exec_, stream_)); }
    void drop_graph() {
        if (graph_exec_) cudaGraphExecDestroy(graph_exec_);
        graph_exec_ = nullptr;
    }
    void read_scalars(double* out) {
        TM_CUDA(cudaMemcpyAsync(h_ms_, ms_, sizeof(double) * MS_COUNT, cudaMemcpyDeviceToHost, stream_));
        TM_CUDA(cudaStreamSynchronize(stream_));
        for (int i = 0; i < MS_COUNT; ++i) out[i] = h_ms_[i];
    }

   private:
    void alloc(double*& p, size_t count) {
        TM_CUDA(cudaMalloc(&p, count * sizeof(double)));
        TM_CUDA(cudaMemset(p, 0, count * sizeof(double)));
        owned_.push_back(p);
    }
    void need_density() const {
        if (!have_density_) throw std::runtime_error("fluid: tm_fluid_set_density must be called first");
    }
    int tri_grid() const { return (int)std::min<size_t>((ntri_ + 127) / 128, (size_t)max_blocks_); }
    int vec_grid() const { return (int)std::min<size_t>((n_ + kVecThreads - 1) / kVecThreads, (size_t)max_blocks_); }
    double read_scalar() {
        TM_CUDA(cudaMemcpyAsync(h_sc_, sc_, sizeof(double), cudaMemcpyDeviceToHost, stream_));
        TM_CUDA(cudaStreamSynchronize(stream_));
        return h_sc_[0];
    }

    int device_, max_blocks_ = 148 * 16;
    FluidGeom g_;
    size_t ntri_, nu_, n1_, n_;
    FluidTables* d_tab_ = nullptr;
    double *Me_ = nullptr, *diag_ = nullptr, *b_ = nullptr, *x_ = nullptr, *xg_ = nullptr, *sc_ = nullptr;
    double* mp_diag_ = nullptr;
    bool mg_ready_ = false;
    double* xprev_ = nullptr;
    bool warm_ = false, have_prev_ = false, last_warm_ = false;
    double *ms_ = nullptr, *h_ms_ = nullptr;  // MINRES scalars on the device / their pinned mirror
    bool dev_scalars_ = false, deterministic_ = false, use_graph_ = false;
    cudaGraphExec_t graph_exec_ = nullptr;
    int check_every_ = 10;
    int precond_mode_ = 0;
    TriMGParams mg_prm_;
    CudaTriMG<6> mg_vel_;
    CudaTriMG<3> mg_prs_;
    double* h_sc_ = nullptr;
    DVec work_[8];
    ReduceScratch rs_{};
    std::vector<double*> owned_;
    cudaStream_t stream_ = nullptr, own_stream_ = nullptr;
    bool have_density_ = false;
};

}  // namespace tmx
