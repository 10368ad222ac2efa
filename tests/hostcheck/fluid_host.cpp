// Host build of topomax_b200/csrc/tm_fluid.cuh: the element arithmetic, the per-triangle work items
// and the MINRES loop of the CUDA fluid path, run serially so that they can be checked against the
// oracle without a GPU.  Test infrastructure only (tests/test_fluid_host.py).
#include <cstring>
#include <type_traits>
#include <vector>

#include "../../topomax_b200/csrc/tm_fluid.cuh"
#include "../../topomax_b200/csrc/tm_trimg.cuh"

namespace {

struct HostFluid {
    tmx::FluidTables T;
    tmx::FluidGeom g;
    size_t ntri, nu, n1, n;
    std::vector<double> Me, diag;

    HostFluid(int nx, int ny, double W, double H, double q, double rmin, double rmax, double visc) {
        T = tmx::fluid_make_tables(W / nx, H / ny);
        g.nx = nx; g.ny = ny; g.Lx = 2 * nx + 1; g.Ly = 2 * ny + 1;
        g.q = q; g.rmin = rmin; g.rmax = rmax; g.viscosity = visc;
        ntri = (size_t)2 * nx * ny;
        nu = (size_t)2 * g.Lx * g.Ly;
        n1 = (size_t)(nx + 1) * (ny + 1);
        n = nu + n1;
    }
    void set_density(const double* rho) {
        Me.assign(21 * ntri, 0.0);
        for (size_t t = 0; t < ntri; ++t) tmx::fluid_body_mass(T, g, rho, Me.data(), ntri, t);
        diag.assign(n, 0.0);
        for (int j = 0; j < g.Ly; ++j)
            for (int i = 0; i < g.Lx; ++i)
                if (i == 0 || j == 0 || i == g.Lx - 1 || j == g.Ly - 1) {
                    diag[2 * ((size_t)j * g.Lx + i)] = 1.0;
                    diag[2 * ((size_t)j * g.Lx + i) + 1] = 1.0;
                }
        for (size_t t = 0; t < ntri; ++t)
            tmx::fluid_body_diag(T, g, Me.data(), ntri, diag.data(), t, tmx::FluidSerialAdd{});
        for (size_t t = 0; t < ntri; ++t) tmx::fluid_body_schur(T, g, diag.data(), t, tmx::FluidSerialAdd{});
    }
    void apply(const double* x, double* y, int mode) {
        std::memset(y, 0, n * sizeof(double));
        for (size_t t = 0; t < ntri; ++t)
            tmx::fluid_body_apply(T, g, Me.data(), ntri, x, y, t, mode, tmx::FluidSerialAdd{});
    }
};

struct HostBackend {
    using Vec = std::vector<double>;
    HostFluid& f;
    std::vector<Vec> ws;
    explicit HostBackend(HostFluid& ff) : f(ff), ws(8, Vec(ff.n, 0.0)) {}
    Vec& work(int i) { return ws[i]; }
    void zero(Vec& a) { std::fill(a.begin(), a.end(), 0.0); }
    void copy(const Vec& a, Vec& b) { b = a; }
    void scale(Vec& a, double s) { for (auto& v : a) v *= s; }
    void apply(const Vec& x, Vec& y) { f.apply(x.data(), y.data(), 0); }
    void precond(const Vec& r, Vec& z) { for (size_t i = 0; i < f.n; ++i) z[i] = r[i] / f.diag[i]; }
    double dot(const Vec& a, const Vec& b) { double s = 0; for (size_t i = 0; i < f.n; ++i) s += a[i] * b[i]; return s; }
    void axpy2(Vec& y, double a, const Vec& p, double b, const Vec& q) {
        for (size_t i = 0; i < f.n; ++i) y[i] += a * p[i] + b * q[i];
    }
    void direction(Vec& wn, const Vec& z, double a3, const Vec& wo, double a2, const Vec& w, double inv_a1, Vec& x,
                   double step) {
        for (size_t i = 0; i < f.n; ++i) {
            wn[i] = (z[i] - a3 * wo[i] - a2 * w[i]) * inv_a1;
            x[i] += step * wn[i];
        }
    }
};

}  // namespace


// ---------------------------------------------------------------------------------------------
// serial multigrid on per-triangle local matrices (tm_trimg.cuh), the back-end of trimg_vcycle
// ---------------------------------------------------------------------------------------------
namespace {

template <int NODES>
struct HostMG {
    using Vec = double*;
    static constexpr int NE = NODES * (NODES + 1) / 2;
    static constexpr int kNodes = NODES;
    std::vector<tmx::TriLevel> geo;
    std::vector<std::vector<double>> Lm, diag, store;
    std::vector<double> lmax_, inv;
    tmx::TriCoarsenTable<NODES> tab = tmx::trimg_make_coarsen_table<NODES>();
    int ncoarse = 0;

    // fine local matrices given; coarsen while both cell counts are even and the level is not tiny
    void build(tmx::TriLevel fine, std::vector<double> L0, int max_levels, int min_cells) {
        geo.assign(1, fine);
        Lm.assign(1, std::move(L0));
        while ((int)geo.size() < max_levels && geo.back().nx % 2 == 0 && geo.back().ny % 2 == 0 &&
               geo.back().nx * geo.back().ny > min_cells) {
            tmx::TriLevel c = geo.back();
            c.nx /= 2;
            c.ny /= 2;
            const size_t ntf = tmx::trimg_num_tri(geo.back()), ntc = tmx::trimg_num_tri(c);
            std::vector<double> Lc((size_t)NE * ntc, 0.0);
            for (size_t t = 0; t < ntc; ++t)
                tmx::trimg_body_coarsen<NODES>(tab, geo.back(), Lm.back().data(), ntf, c, Lc.data(), ntc, t);
            geo.push_back(c);
            Lm.push_back(std::move(Lc));
        }
        const int L = (int)geo.size();
        diag.resize(L);
        store.resize(L);
        lmax_.assign(L, 0.0);
        for (int l = 0; l < L; ++l) {
            const size_t n = size(l);
            diag[l].assign(n, 0.0);
            if (NODES == 6 && geo[l].fixed_boundary) {
                const int Lx = 2 * geo[l].nx + 1, Ly = 2 * geo[l].ny + 1;
                for (int j = 0; j < Ly; ++j)
                    for (int i = 0; i < Lx; ++i)
                        if (i == 0 || j == 0 || i == Lx - 1 || j == Ly - 1)
                            for (int c = 0; c < geo[l].ncomp; ++c) diag[l][((size_t)j * Lx + i) * geo[l].ncomp + c] = 1.0;
            }
            for (size_t t = 0; t < tmx::trimg_num_tri(geo[l]); ++t)
                tmx::trimg_body_diag<NODES>(geo[l], Lm[l].data(), tmx::trimg_num_tri(geo[l]), diag[l].data(), t,
                                            tmx::FluidSerialAdd{});
            store[l].assign((size_t)tmx::TRIMG_NVEC * n, 0.0);
            // lambda_max(D^-1 A) by power iteration
            std::vector<double> v(n), w(n);
            for (size_t i = 0; i < n; ++i) v[i] = 0.25 + (double)((i * 2654435761u) % 1000) / 1000.0;
            double lam = 1.0;
            for (int it = 0; it < 30; ++it) {
                apply(l, v.data(), w.data());
                double nrm = 0.0;
                for (size_t i = 0; i < n; ++i) { w[i] /= diag[l][i]; nrm += w[i] * w[i]; }
                nrm = std::sqrt(nrm);
                double vv = 0.0;
                for (size_t i = 0; i < n; ++i) vv += v[i] * v[i];
                lam = nrm / std::sqrt(vv);
                for (size_t i = 0; i < n; ++i) v[i] = w[i] / nrm;
            }
            lmax_[l] = 1.1 * lam;
        }
        // explicit inverse of the coarsest scalar operator (the same for every component)
        inv = tmx::trimg_coarse_inverse<NODES>(geo[L - 1], Lm[L - 1].data(), ncoarse);
    }
    size_t size(int l) const { return tmx::trimg_num_nodes<NODES>(geo[l]) * geo[l].ncomp; }
    void apply(int l, const double* x, double* y) {
        const size_t n = size(l), nt = tmx::trimg_num_tri(geo[l]);
        std::fill(y, y + n, 0.0);
        for (size_t t = 0; t < nt; ++t)
            tmx::trimg_body_apply<NODES>(geo[l], Lm[l].data(), nt, x, y, t, tmx::FluidSerialAdd{});
    }
    // ---- back-end interface of trimg_vcycle
    int levels() const { return (int)geo.size(); }
    double lmax(int l) const { return lmax_[l]; }
    Vec vec(int l, int which) { return store[l].data() + (size_t)which * size(l); }
    void residual(int l, Vec b, Vec x, Vec r) {
        apply(l, x, r);
        for (size_t i = 0; i < size(l); ++i) r[i] = b[i] - r[i];
    }
    void cheb_first(int l, Vec b, Vec d, Vec x, double s) {
        for (size_t i = 0; i < size(l); ++i) d[i] = x[i] = s * b[i] / diag[l][i];
    }
    void cheb_next(int l, Vec r, Vec d, Vec x, double c1, double c2) {
        for (size_t i = 0; i < size(l); ++i) {
            d[i] = c1 * d[i] + c2 * r[i] / diag[l][i];
            x[i] += d[i];
        }
    }
    void restrict_to(int l, Vec rf, Vec bc) {
        std::fill(bc, bc + size(l + 1), 0.0);
        for (size_t nf = 0; nf < tmx::trimg_num_nodes<NODES>(geo[l]); ++nf)
            tmx::trimg_body_restrict<NODES>(geo[l], geo[l + 1], rf, bc, nf, tmx::FluidSerialAdd{});
    }
    void prolong_add(int l, Vec xc, Vec xf) {
        for (size_t nf = 0; nf < tmx::trimg_num_nodes<NODES>(geo[l]); ++nf)
            tmx::trimg_body_prolong_add<NODES>(geo[l], geo[l + 1], xc, xf, nf);
    }
    void coarse_solve(Vec b, Vec x) {
        const int nc = geo.back().ncomp;
        for (int c = 0; c < nc; ++c)
            for (int i = 0; i < ncoarse; ++i) {
                double s = 0.0;
                for (int j = 0; j < ncoarse; ++j) s += inv[(size_t)i * ncoarse + j] * b[(size_t)j * nc + c];
                x[(size_t)i * nc + c] = s;
            }
    }
    void add(int l, Vec e, Vec x) {
        for (size_t i = 0; i < size(l); ++i) x[i] += e[i];
    }
    // z = V-cycle(r)
    void precondition(const double* r, double* z, const tmx::TriMGParams& prm) {
        std::copy(r, r + size(0), vec(0, tmx::TRIMG_B));
        tmx::trimg_vcycle(*this, 0, prm);
        std::copy(vec(0, tmx::TRIMG_X), vec(0, tmx::TRIMG_X) + size(0), z);
    }
};

// fine local matrices of the two hierarchies of the fluid preconditioner
std::vector<double> velocity_local_matrices(const HostFluid& f) {
    std::vector<double> L(21 * f.ntri);
    for (size_t t = 0; t < f.ntri; ++t) {
        const int ty = (int)(t & 1);
        int k = 0;
        for (int i = 0; i < 6; ++i)
            for (int j = i; j < 6; ++j, ++k) L[(size_t)k * f.ntri + t] = f.Me[(size_t)k * f.ntri + t] + f.T.Kref[ty][i][j];
    }
    return L;
}
std::vector<double> darcy_local_matrices(const HostFluid& f, const double* rho) {
    std::vector<double> L(6 * f.ntri);
    for (size_t t = 0; t < f.ntri; ++t) tmx::fluid_body_darcy(f.T, f.g, rho, L.data(), f.ntri, t);
    return L;
}

struct HostBackendMG : HostBackend {
    HostMG<6>& vel;
    HostMG<3>& prs;
    std::vector<double> mp_diag;
    tmx::TriMGParams prm;
    HostBackendMG(HostFluid& ff, HostMG<6>& v, HostMG<3>& p) : HostBackend(ff), vel(v), prs(p), mp_diag(ff.n1, 0.0) {
        for (size_t t = 0; t < ff.ntri; ++t) tmx::fluid_body_pmass_diag(ff.T, ff.g, mp_diag.data(), t, tmx::FluidSerialAdd{});
    }
    void precond(const Vec& r, Vec& z) {
        vel.precondition(r.data(), z.data(), prm);
        prs.precondition(r.data() + f.nu, z.data() + f.nu, prm);
        for (size_t i = 0; i < f.n1; ++i) z[f.nu + i] += r[f.nu + i] / mp_diag[i];
    }
};

}  // namespace

extern "C" {

void hc_fluid_tables(double hx, double hy, double* Kref72, double* Dloc72) {
    const tmx::FluidTables T = tmx::fluid_make_tables(hx, hy);
    std::memcpy(Kref72, T.Kref, sizeof(T.Kref));
    std::memcpy(Dloc72, T.Dloc, sizeof(T.Dloc));
}

// y = Op x  (mode 0: homogeneous operator, boundary rows zero; mode 1: lifting of boundary values)
void hc_fluid_apply(int nx, int ny, double W, double H, double q, double rmin, double rmax, double visc,
                    const double* rho, const double* x, double* y, int mode, double* diag_out) {
    HostFluid f(nx, ny, W, H, q, rmin, rmax, visc);
    f.set_density(rho);
    f.apply(x, y, mode);
    if (diag_out) std::memcpy(diag_out, f.diag.data(), f.n * sizeof(double));
}

// the whole state solve as the CUDA driver does it: lift, project, MINRES, add the boundary values
int hc_fluid_solve(int nx, int ny, double W, double H, double q, double rmin, double rmax, double visc,
                   const double* rho, const double* g_boundary, double rtol, int maxit, double* up,
                   double* relres) {
    HostFluid f(nx, ny, W, H, q, rmin, rmax, visc);
    f.set_density(rho);
    std::vector<double> xg(f.n, 0.0), b(f.n, 0.0), x(f.n, 0.0);
    std::memcpy(xg.data(), g_boundary, f.nu * sizeof(double));
    f.apply(xg.data(), b.data(), 1);
    double mean = 0.0;
    for (size_t i = f.nu; i < f.n; ++i) mean += b[i];
    mean /= (double)f.n1;
    for (size_t i = 0; i < f.n; ++i) b[i] = -(b[i] - (i >= f.nu ? mean : 0.0));
    HostBackend bk(f);
    const tmx::MinresResult r = tmx::fluid_minres(bk, b, x, rtol, maxit);
    for (size_t i = 0; i < f.n; ++i) up[i] = x[i] + (i < f.nu ? g_boundary[i] : 0.0);
    *relres = r.relres;
    return r.converged ? r.iterations : -r.iterations;
}

// level operators of the two hierarchies (kind 0: velocity block M_r + K with Dirichlet boundary,
// 2 components; kind 1: P1 Darcy Laplacian) and their transfers, for the Galerkin checks
int hc_trimg_level_apply(int nx, int ny, double W, double H, double q, double rmin, double rmax, double visc,
                         const double* rho, int kind, int level, const double* x, double* y) {
    HostFluid f(nx, ny, W, H, q, rmin, rmax, visc);
    f.set_density(rho);
    if (kind == 0) {
        HostMG<6> mg;
        mg.build(tmx::TriLevel{nx, ny, 2, 1}, velocity_local_matrices(f), 16, 1);
        if (level >= mg.levels()) return -mg.levels();
        mg.apply(level, x, y);
        return mg.levels();
    }
    HostMG<3> mg;
    mg.build(tmx::TriLevel{nx, ny, 1, 0}, darcy_local_matrices(f, rho), 16, 1);
    if (level >= mg.levels()) return -mg.levels();
    mg.apply(level, x, y);
    return mg.levels();
}
// the gather form of the level operator (every node written once) against the scatter form
double hc_trimg_gather_vs_scatter(int nx, int ny, double W, double H, double q, double rmin, double rmax,
                                  double visc, const double* rho, int kind, int level, const double* x) {
    HostFluid f(nx, ny, W, H, q, rmin, rmax, visc);
    f.set_density(rho);
    double worst = 0.0;
    auto compare = [&](auto& mg) {
        if (level >= mg.levels()) { worst = -1.0; return; }
        const size_t n = mg.size(level);
        std::vector<double> ys(n), yg(n, 123.0);
        mg.apply(level, x, ys.data());
        const size_t nodes = n / mg.geo[level].ncomp;
        for (size_t k = 0; k < nodes; ++k)
            tmx::trimg_body_apply_gather<std::remove_reference_t<decltype(mg)>::kNodes>(
                mg.geo[level], mg.Lm[level].data(), tmx::trimg_num_tri(mg.geo[level]), x, yg.data(), k);
        double scale = 0.0;
        for (size_t i = 0; i < n; ++i) scale = std::max(scale, std::fabs(ys[i]));
        for (size_t i = 0; i < n; ++i) worst = std::max(worst, std::fabs(ys[i] - yg[i]) / (scale > 0 ? scale : 1.0));
    };
    if (kind == 0) {
        HostMG<6> mg;
        mg.build(tmx::TriLevel{nx, ny, 2, 1}, velocity_local_matrices(f), 16, 1);
        compare(mg);
    } else {
        HostMG<3> mg;
        mg.build(tmx::TriLevel{nx, ny, 1, 0}, darcy_local_matrices(f, rho), 16, 1);
        compare(mg);
    }
    return worst;
}
void hc_trimg_prolong(int kind, int nxf, int nyf, const double* xc, double* xf) {
    if (kind == 0) {
        const tmx::TriLevel gf{nxf, nyf, 2, 1}, gc{nxf / 2, nyf / 2, 2, 1};
        for (size_t n = 0; n < tmx::trimg_num_nodes<6>(gf); ++n) tmx::trimg_body_prolong_add<6>(gf, gc, xc, xf, n);
    } else {
        const tmx::TriLevel gf{nxf, nyf, 1, 0}, gc{nxf / 2, nyf / 2, 1, 0};
        for (size_t n = 0; n < tmx::trimg_num_nodes<3>(gf); ++n) tmx::trimg_body_prolong_add<3>(gf, gc, xc, xf, n);
    }
}
void hc_trimg_restrict(int kind, int nxf, int nyf, const double* rf, double* rc) {
    if (kind == 0) {
        const tmx::TriLevel gf{nxf, nyf, 2, 1}, gc{nxf / 2, nyf / 2, 2, 1};
        for (size_t n = 0; n < tmx::trimg_num_nodes<6>(gf); ++n)
            tmx::trimg_body_restrict<6>(gf, gc, rf, rc, n, tmx::FluidSerialAdd{});
    } else {
        const tmx::TriLevel gf{nxf, nyf, 1, 0}, gc{nxf / 2, nyf / 2, 1, 0};
        for (size_t n = 0; n < tmx::trimg_num_nodes<3>(gf); ++n)
            tmx::trimg_body_restrict<3>(gf, gc, rf, rc, n, tmx::FluidSerialAdd{});
    }
}

// z = V-cycle(r) of hierarchy `kind` (0 velocity, 1 pressure Darcy Laplacian)
void hc_trimg_vcycle(int nx, int ny, double W, double H, double q, double rmin, double rmax, double visc,
                     const double* rho, int kind, const double* r, double* z) {
    HostFluid f(nx, ny, W, H, q, rmin, rmax, visc);
    f.set_density(rho);
    tmx::TriMGParams prm;
    if (kind == 0) {
        HostMG<6> mg;
        mg.build(tmx::TriLevel{nx, ny, 2, 1}, velocity_local_matrices(f), 16, 4);
        mg.precondition(r, z, prm);
    } else {
        HostMG<3> mg;
        mg.build(tmx::TriLevel{nx, ny, 1, 0}, darcy_local_matrices(f, rho), 16, 4);
        mg.precondition(r, z, prm);
    }
}

// the state solve with the multigrid preconditioner (velocity V-cycle | M_p^-1 + Darcy V-cycle)
int hc_fluid_solve_mg(int nx, int ny, double W, double H, double q, double rmin, double rmax, double visc,
                      const double* rho, const double* g_boundary, double rtol, int maxit, double* up,
                      double* relres) {
    HostFluid f(nx, ny, W, H, q, rmin, rmax, visc);
    f.set_density(rho);
    HostMG<6> vel;
    vel.build(tmx::TriLevel{nx, ny, 2, 1}, velocity_local_matrices(f), 16, 4);
    HostMG<3> prs;
    prs.build(tmx::TriLevel{nx, ny, 1, 0}, darcy_local_matrices(f, rho), 16, 4);
    std::vector<double> xg(f.n, 0.0), b(f.n, 0.0), x(f.n, 0.0);
    std::memcpy(xg.data(), g_boundary, f.nu * sizeof(double));
    f.apply(xg.data(), b.data(), 1);
    double mean = 0.0;
    for (size_t i = f.nu; i < f.n; ++i) mean += b[i];
    mean /= (double)f.n1;
    for (size_t i = 0; i < f.n; ++i) b[i] = -(b[i] - (i >= f.nu ? mean : 0.0));
    HostBackendMG bk(f, vel, prs);
    const tmx::MinresResult r = tmx::fluid_minres(bk, b, x, rtol, maxit);
    for (size_t i = 0; i < f.n; ++i) up[i] = x[i] + (i < f.nu ? g_boundary[i] : 0.0);
    *relres = r.relres;
    return r.converged ? r.iterations : -r.iterations;
}

double hc_fluid_objective(int nx, int ny, double W, double H, double q, double rmin, double rmax, double visc,
                          const double* rho, const double* u) {
    HostFluid f(nx, ny, W, H, q, rmin, rmax, visc);
    f.set_density(rho);
    double s = 0.0;
    for (size_t t = 0; t < f.ntri; ++t) s += tmx::fluid_body_objective(f.T, f.g, f.Me.data(), f.ntri, u, t);
    return s;
}

void hc_fluid_sens(int nx, int ny, double W, double H, double q, double rmin, double rmax, double visc,
                   const double* rho, const double* u, double* out) {
    HostFluid f(nx, ny, W, H, q, rmin, rmax, visc);
    std::memset(out, 0, f.n1 * sizeof(double));
    for (size_t t = 0; t < f.ntri; ++t) tmx::fluid_body_sens(f.T, f.g, rho, u, out, t, tmx::FluidSerialAdd{});
}
}
