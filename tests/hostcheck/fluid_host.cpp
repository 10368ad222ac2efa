// Host build of topomax_b200/csrc/tm_fluid.cuh: the element arithmetic, the per-triangle work items
// and the MINRES loop of the CUDA fluid path, run serially so that they can be checked against the
// oracle without a GPU.  Test infrastructure only (tests/test_fluid_host.py).
#include <cstring>
#include <vector>

#include "../../topomax_b200/csrc/tm_fluid.cuh"

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
