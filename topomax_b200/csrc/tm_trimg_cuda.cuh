// CUDA back-end of the per-triangle multigrid (tm_trimg.cuh): launches around the shared work items
// and the back-end interface of trimg_vcycle.  One instance per hierarchy (NODES = 6: velocity block,
// NODES = 3: pressure "Darcy" Laplacian).  Opt-in preconditioner of the fluid solver.
#pragma once

#include <stdexcept>
#include <vector>

#ifdef TM_HOST_SHIM  // serial CPU build for the CPU tests (tests/hostcheck/cuda_host_shim.h)
#include "cuda_host_shim.h"
#else
#include <cuda_runtime.h>

#include "tm_common.cuh"
#include "tm_vec.cuh"
#endif
#include "tm_trimg.cuh"

namespace tmx {

struct TriAtomicAdd {
    __host__ __device__ __forceinline__ void operator()(double* p, double v) const {
#ifdef __CUDA_ARCH__
        atomicAdd(p, v);
#else
        *p += v;
#endif
    }
};

template <int NODES>
__global__ void __launch_bounds__(128) trimg_apply_kernel(TriLevel g, const double* __restrict__ Lm, size_t ntri,
                                                          const double* __restrict__ x, double* y) {
    TM_GRID_STRIDE(tid, ntri) trimg_body_apply<NODES>(g, Lm, ntri, x, y, tid, TriAtomicAdd{});
}
template <int NODES>
__global__ void __launch_bounds__(128) trimg_diag_kernel(TriLevel g, const double* __restrict__ Lm, size_t ntri,
                                                         double* diag) {
    TM_GRID_STRIDE(tid, ntri) trimg_body_diag<NODES>(g, Lm, ntri, diag, tid, TriAtomicAdd{});
}
// diag = 1 on Dirichlet nodes, 0 elsewhere (before the scatter)
template <int NODES>
__global__ void trimg_diag_init_kernel(TriLevel g, double* diag, size_t n) {
    const int Lx = 2 * g.nx + 1, Ly = 2 * g.ny + 1;
    TM_GRID_STRIDE(k, n) {
        double v = 0.0;
        if (NODES == 6 && g.fixed_boundary) {
            const size_t node = k / g.ncomp;
            const int j = (int)(node / Lx), i = (int)(node - (size_t)j * Lx);
            if (i == 0 || j == 0 || i == Lx - 1 || j == Ly - 1) v = 1.0;
        }
        diag[k] = v;
    }
}
template <int NODES>
__global__ void __launch_bounds__(128) trimg_coarsen_kernel(const TriCoarsenTable<NODES>* __restrict__ tab, TriLevel gf,
                                                            const double* __restrict__ Lf, size_t ntri_f, TriLevel gc,
                                                            double* __restrict__ Lc, size_t ntri_c) {
    TM_GRID_STRIDE(tid, ntri_c) trimg_body_coarsen<NODES>(*tab, gf, Lf, ntri_f, gc, Lc, ntri_c, tid);
}
template <int NODES>
__global__ void trimg_prolong_kernel(TriLevel gf, TriLevel gc, const double* __restrict__ xc, double* xf, size_t nodes) {
    TM_GRID_STRIDE(nf, nodes) trimg_body_prolong_add<NODES>(gf, gc, xc, xf, nf);
}
template <int NODES>
__global__ void trimg_restrict_kernel(TriLevel gf, TriLevel gc, const double* __restrict__ rf, double* rc,
                                      size_t nodes) {
    TM_GRID_STRIDE(nf, nodes) trimg_body_restrict<NODES>(gf, gc, rf, rc, nf, TriAtomicAdd{});
}
// gather forms (deterministic mode): one work item per output node, no atomics, no memset
template <int NODES>
__global__ void __launch_bounds__(128) trimg_apply_gather_kernel(TriLevel g, const double* __restrict__ Lm, size_t ntri,
                                                                 const double* __restrict__ x, double* __restrict__ y,
                                                                 size_t nodes) {
    TM_GRID_STRIDE(n, nodes) trimg_body_apply_gather<NODES>(g, Lm, ntri, x, y, n);
}
template <int NODES>
__global__ void trimg_diag_gather_kernel(TriLevel g, const double* __restrict__ Lm, size_t ntri, double* diag,
                                         size_t nodes) {
    TM_GRID_STRIDE(n, nodes) trimg_body_diag_gather<NODES>(g, Lm, ntri, diag, n);
}
template <int NODES>
__global__ void trimg_restrict_gather_kernel(TriLevel gf, TriLevel gc, const double* __restrict__ rf,
                                             double* __restrict__ rc, size_t cnodes) {
    TM_GRID_STRIDE(nc, cnodes) trimg_body_restrict_gather<NODES>(gf, gc, rf, rc, nc);
}
// r = b - r   (r holds A x)
__global__ void trimg_residual_kernel(size_t n, const double* __restrict__ b, double* r) {
    TM_GRID_STRIDE(i, n) r[i] = b[i] - r[i];
}
__global__ void trimg_cheb_first_kernel(size_t n, const double* __restrict__ diag, const double* __restrict__ b,
                                        double* __restrict__ d, double* __restrict__ x, double s) {
    TM_GRID_STRIDE(i, n) {
        const double v = s * b[i] / diag[i];
        d[i] = v;
        x[i] = v;
    }
}
__global__ void trimg_cheb_next_kernel(size_t n, const double* __restrict__ diag, const double* __restrict__ r,
                                       double* __restrict__ d, double* __restrict__ x, double c1, double c2) {
    TM_GRID_STRIDE(i, n) {
        const double v = c1 * d[i] + c2 * r[i] / diag[i];
        d[i] = v;
        x[i] += v;
    }
}
__global__ void trimg_add_kernel(size_t n, const double* __restrict__ e, double* x) { TM_GRID_STRIDE(i, n) x[i] += e[i]; }
// x[i*nc + c] = sum_j inv[i][j] b[j*nc + c]   (one thread per output entry; the level is tiny)
__global__ void trimg_coarse_solve_kernel(int n, int nc, const double* __restrict__ inv, const double* __restrict__ b,
                                          double* __restrict__ x) {
    TM_GRID_STRIDE(k, (size_t)n * nc) {
        const int i = (int)(k / nc), c = (int)(k - (size_t)i * nc);
        double s = 0.0;
        for (int j = 0; j < n; ++j) s += inv[(size_t)i * n + j] * b[(size_t)j * nc + c];
        x[k] = s;
    }
}
// w = (A v) / diag in place of w = A v
__global__ void trimg_scale_diag_kernel(size_t n, const double* __restrict__ diag, double* w) {
    TM_GRID_STRIDE(i, n) w[i] /= diag[i];
}
__global__ void trimg_fill_kernel(size_t n, double* v) {
    TM_GRID_STRIDE(i, n) v[i] = 0.25 + (double)((i * 2654435761ull) % 1000ull) / 1000.0;
}
__global__ void trimg_scale_kernel(size_t n, const double* __restrict__ w, double* v, double s) {
    TM_GRID_STRIDE(i, n) v[i] = s * w[i];
}

template <int NODES>
class CudaTriMG {
   public:
    using Vec = double*;
    static constexpr int NE = NODES * (NODES + 1) / 2;
    static constexpr int kMaxCoarse = 2048;  // scalar nodes of the coarsest level (explicit inverse)

    CudaTriMG() = default;
    CudaTriMG(const CudaTriMG&) = delete;
    CudaTriMG& operator=(const CudaTriMG&) = delete;
    ~CudaTriMG() { release(); }

    // plan the levels and allocate; fine local matrices live in level_matrices(0), filled by the caller
    // level list of a mesh; throws when the coarsest level would be too large for the dense inverse
    static std::vector<TriLevel> level_list(TriLevel fine, int min_cells = 4) {
        std::vector<TriLevel> geo(1, fine);
        while (geo.back().nx % 2 == 0 && geo.back().ny % 2 == 0 && geo.back().nx * geo.back().ny > min_cells) {
            TriLevel c = geo.back();
            c.nx /= 2;
            c.ny /= 2;
            geo.push_back(c);
        }
        const TriLevel one{geo.back().nx, geo.back().ny, 1, geo.back().fixed_boundary};
        if (trimg_num_nodes<NODES>(one) > (size_t)kMaxCoarse)
            throw std::runtime_error("fluid multigrid: the mesh cannot be coarsened far enough (cell counts need "
                                     "more factors of two); use the diagonal preconditioner");
        return geo;
    }
    // Exception safe: the level list is validated before any state changes, and a failure while
    // allocating leaves the object un-planned (planned() == false), never half built.
    void plan(TriLevel fine, int max_blocks, int min_cells = 4) {
        std::vector<TriLevel> geo = level_list(fine, min_cells);  // may throw: nothing touched yet
        release();
        try {
            plan_validated(std::move(geo), max_blocks);
        } catch (...) {
            release();
            throw;
        }
    }
    void plan_validated(std::vector<TriLevel> geo, int max_blocks) {
        max_blocks_ = max_blocks;
        geo_ = std::move(geo);
        const TriLevel one{geo_.back().nx, geo_.back().ny, 1, geo_.back().fixed_boundary};
        const int L = (int)geo_.size();
        Lm_.assign(L, nullptr);
        diag_.assign(L, nullptr);
        store_.assign(L, nullptr);
        lmax_.assign(L, 1.0);
        for (int l = 0; l < L; ++l) {
            alloc(Lm_[l], (size_t)NE * trimg_num_tri(geo_[l]));
            alloc(diag_[l], size(l));
            alloc(store_[l], (size_t)TRIMG_NVEC * size(l));
        }
        ncoarse_ = (int)trimg_num_nodes<NODES>(one);
        alloc(inv_, (size_t)ncoarse_ * ncoarse_);
        const TriCoarsenTable<NODES> tab = trimg_make_coarsen_table<NODES>();
        TM_CUDA(cudaMalloc(&tab_, sizeof(tab)));
        TM_CUDA(cudaMemcpy(tab_, &tab, sizeof(tab), cudaMemcpyHostToDevice));
    }
    bool planned() const { return !geo_.empty(); }
    void set_deterministic(bool on) { deterministic_ = on; }
    double* level_matrices(int l) { return Lm_[l]; }

    // Galerkin coarse matrices, diagonals, smoothing bounds and the coarsest inverse from level 0
    void setup(cudaStream_t st, ReduceScratch rs, double* d_scalar, double* h_scalar) {
        stream_ = st;
        const int L = levels();
        for (int l = 0; l + 1 < L; ++l) {
            const size_t ntc = trimg_num_tri(geo_[l + 1]);
            TM_LAUNCH(trimg_coarsen_kernel<NODES>, grid(ntc, 128), 128, st)(tab_, geo_[l], Lm_[l], trimg_num_tri(geo_[l]),
                                                                       geo_[l + 1], Lm_[l + 1], ntc);
            TM_CHECK_LAUNCH();
        }
        for (int l = 0; l < L; ++l) {
            const size_t n = size(l), nt = trimg_num_tri(geo_[l]);
            if (deterministic_) {
                const size_t nodes = trimg_num_nodes<NODES>(geo_[l]);
                TM_LAUNCH(trimg_diag_gather_kernel<NODES>, grid(nodes, 256), 256, st)(geo_[l], Lm_[l], nt, diag_[l], nodes);
                TM_CHECK_LAUNCH();
                continue;
            }
            TM_LAUNCH(trimg_diag_init_kernel<NODES>, grid(n, 256), 256, st)(geo_[l], diag_[l], n);
            TM_CHECK_LAUNCH();
            TM_LAUNCH(trimg_diag_kernel<NODES>, grid(nt, 128), 128, st)(geo_[l], Lm_[l], nt, diag_[l]);
            TM_CHECK_LAUNCH();
        }
        // lambda_max(D^-1 A) per level by power iteration (30 steps, like the host check)
        for (int l = 0; l + 1 < L; ++l) {
            const size_t n = size(l);
            double *v = vec(l, TRIMG_D), *w = vec(l, TRIMG_E);
            TM_LAUNCH(trimg_fill_kernel, grid(n, 256), 256, st)(n, v);
            TM_CHECK_LAUNCH();
            double lam = 1.0, vv = 0.0;
            TM_LAUNCH(dot_kernel<double>, grid(n, kVecThreads), kVecThreads, st)(n, v, v, rs, d_scalar);
            TM_CHECK_LAUNCH();
            vv = read(d_scalar, h_scalar);
            for (int it = 0; it < 30; ++it) {
                apply(l, v, w);
                TM_LAUNCH(trimg_scale_diag_kernel, grid(n, 256), 256, st)(n, diag_[l], w);
                TM_CHECK_LAUNCH();
                TM_LAUNCH(dot_kernel<double>, grid(n, kVecThreads), kVecThreads, st)(n, w, w, rs, d_scalar);
                TM_CHECK_LAUNCH();
                const double ww = read(d_scalar, h_scalar);
                if (!(ww > 0.0) || !(vv > 0.0)) break;
                lam = std::sqrt(ww / vv);
                TM_LAUNCH(trimg_scale_kernel, grid(n, 256), 256, st)(n, w, v, 1.0 / std::sqrt(ww));
                TM_CHECK_LAUNCH();
                vv = 1.0;
            }
            lmax_[l] = 1.1 * lam;
        }
        // coarsest level: explicit inverse assembled on the host
        const size_t cnt = (size_t)NE * trimg_num_tri(geo_[L - 1]);
        std::vector<double> hL(cnt);
        TM_CUDA(cudaMemcpyAsync(hL.data(), Lm_[L - 1], cnt * sizeof(double), cudaMemcpyDeviceToHost, st));
        TM_CUDA(cudaStreamSynchronize(st));
        int n = 0;
        const std::vector<double> inv = trimg_coarse_inverse<NODES>(geo_[L - 1], hL.data(), n);
        TM_CUDA(cudaMemcpyAsync(inv_, inv.data(), inv.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        TM_CUDA(cudaStreamSynchronize(st));  // `inv` is a pageable temporary
    }

    // z = V-cycle(r) on level-0 vectors
    void precondition(const double* r, double* z, const TriMGParams& prm) {
        TM_CUDA(cudaMemcpyAsync(vec(0, TRIMG_B), r, size(0) * sizeof(double), cudaMemcpyDeviceToDevice, stream_));
        trimg_vcycle(*this, 0, prm);
        TM_CUDA(cudaMemcpyAsync(z, vec(0, TRIMG_X), size(0) * sizeof(double), cudaMemcpyDeviceToDevice, stream_));
    }

    // ---- back-end interface of trimg_vcycle
    int levels() const { return (int)geo_.size(); }
    double lmax(int l) const { return lmax_[l]; }
    size_t size(int l) const { return trimg_num_nodes<NODES>(geo_[l]) * geo_[l].ncomp; }
    Vec vec(int l, int which) { return store_[l] + (size_t)which * size(l); }
    void apply(int l, const double* x, double* y) {
        const size_t nt = trimg_num_tri(geo_[l]);
        if (deterministic_) {
            const size_t nodes = trimg_num_nodes<NODES>(geo_[l]);
            TM_LAUNCH(trimg_apply_gather_kernel<NODES>, grid(nodes, 128), 128, stream_)(geo_[l], Lm_[l], nt, x, y, nodes);
            TM_CHECK_LAUNCH();
            return;
        }
        TM_CUDA(cudaMemsetAsync(y, 0, size(l) * sizeof(double), stream_));
        TM_LAUNCH(trimg_apply_kernel<NODES>, grid(nt, 128), 128, stream_)(geo_[l], Lm_[l], nt, x, y);
        TM_CHECK_LAUNCH();
    }
    void residual(int l, Vec b, Vec x, Vec r) {
        apply(l, x, r);
        TM_LAUNCH(trimg_residual_kernel, grid(size(l), 256), 256, stream_)(size(l), b, r);
        TM_CHECK_LAUNCH();
    }
    void cheb_first(int l, Vec b, Vec d, Vec x, double s) {
        TM_LAUNCH(trimg_cheb_first_kernel, grid(size(l), 256), 256, stream_)(size(l), diag_[l], b, d, x, s);
        TM_CHECK_LAUNCH();
    }
    void cheb_next(int l, Vec r, Vec d, Vec x, double c1, double c2) {
        TM_LAUNCH(trimg_cheb_next_kernel, grid(size(l), 256), 256, stream_)(size(l), diag_[l], r, d, x, c1, c2);
        TM_CHECK_LAUNCH();
    }
    void restrict_to(int l, Vec rf, Vec bc) {
        if (deterministic_) {
            const size_t cnodes = trimg_num_nodes<NODES>(geo_[l + 1]);
            TM_LAUNCH(trimg_restrict_gather_kernel<NODES>, grid(cnodes, 128), 128, stream_)(geo_[l], geo_[l + 1], rf, bc,
                                                                                           cnodes);
            TM_CHECK_LAUNCH();
            return;
        }
        const size_t nodes = trimg_num_nodes<NODES>(geo_[l]);
        TM_CUDA(cudaMemsetAsync(bc, 0, size(l + 1) * sizeof(double), stream_));
        TM_LAUNCH(trimg_restrict_kernel<NODES>, grid(nodes, 256), 256, stream_)(geo_[l], geo_[l + 1], rf, bc, nodes);
        TM_CHECK_LAUNCH();
    }
    void prolong_add(int l, Vec xc, Vec xf) {
        const size_t nodes = trimg_num_nodes<NODES>(geo_[l]);
        TM_LAUNCH(trimg_prolong_kernel<NODES>, grid(nodes, 256), 256, stream_)(geo_[l], geo_[l + 1], xc, xf, nodes);
        TM_CHECK_LAUNCH();
    }
    void coarse_solve(Vec b, Vec x) {
        const int nc = geo_.back().ncomp;
        TM_LAUNCH(trimg_coarse_solve_kernel, grid((size_t)ncoarse_ * nc, 128), 128, stream_)(ncoarse_, nc, inv_, b, x);
        TM_CHECK_LAUNCH();
    }
    void add(int l, Vec e, Vec x) {
        TM_LAUNCH(trimg_add_kernel, grid(size(l), 256), 256, stream_)(size(l), e, x);
        TM_CHECK_LAUNCH();
    }

   private:
    int grid(size_t n, int threads) const {
        const size_t b = (n + threads - 1) / threads;
        return (int)std::max<size_t>(1, std::min<size_t>(b, (size_t)max_blocks_));
    }
    void alloc(double*& p, size_t count) {
        TM_CUDA(cudaMalloc(&p, count * sizeof(double)));
        TM_CUDA(cudaMemset(p, 0, count * sizeof(double)));
        owned_.push_back(p);
    }
    double read(double* d_scalar, double* h_scalar) {
        TM_CUDA(cudaMemcpyAsync(h_scalar, d_scalar, sizeof(double), cudaMemcpyDeviceToHost, stream_));
        TM_CUDA(cudaStreamSynchronize(stream_));
        return h_scalar[0];
    }
    void release() {
        for (double* p : owned_) cudaFree(p);
        owned_.clear();
        if (tab_) cudaFree(tab_);
        tab_ = nullptr;
        inv_ = nullptr;
        Lm_.clear();
        diag_.clear();
        store_.clear();
        geo_.clear();
    }

    std::vector<TriLevel> geo_;
    std::vector<double*> Lm_, diag_, store_, owned_;
    std::vector<double> lmax_;
    double* inv_ = nullptr;
    TriCoarsenTable<NODES>* tab_ = nullptr;
    int ncoarse_ = 0, max_blocks_ = 148 * 16;
    bool deterministic_ = false;
    cudaStream_t stream_ = nullptr;
};

}  // namespace tmx
