"""CPU study (scipy, not a GPU number): PCG iteration counts of the elasticity multigrid for smoother / cycle
variants, on the oracle's matrices with the library's hierarchy (nested P2 spaces, Galerkin coarse operators,
Chebyshev-Jacobi smoothing 1 step fine / 3 coarse on [lmax/30, 1.1 lmax], V-cycle, exact coarsest solve).

    python tools/studies/elast_smoother_study.py [N_per_unit=32] [md_iterations=25]

Variants: the Chebyshev polynomial (first kind on an interval = what the library runs; fourth kind and the
optimised fourth kind of Lottes, "Optimal polynomial smoothers for multigrid V-cycles", 2022: they need lmax only),
2x2 node-block Jacobi, degrees, W-cycle on the coarse levels.
"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.fem_oracle import StructuredMesh  # noqa: E402
from oracle.md_oracle import OracleSolver, expit, logit  # noqa: E402


def prolongation(nxc, nyc):
    """Scalar P2 interpolation from the (nxc, nyc) mesh to the uniformly refined (2nxc, 2nyc) mesh."""
    nxf, nyf = 2 * nxc, 2 * nyc
    Lxf, Lyf, Lxc = 2 * nxf + 1, 2 * nyf + 1, 2 * nxc + 1
    i, j = np.meshgrid(np.arange(Lxf), np.arange(Lyf), indexing="xy")
    i, j = i.ravel(), j.ravel()
    cx = np.minimum(i // 4, nxc - 1)
    cy = np.minimum(j // 4, nyc - 1)
    s = (i - 4 * cx) / 4.0
    t = (j - 4 * cy) / 4.0
    lowerA = s >= t
    rows, cols, vals = [], [], []
    for isA in (True, False):
        m = lowerA if isA else ~lowerA
        ss, tt, ccx, ccy = s[m], t[m], cx[m], cy[m]
        fine = (j[m] * Lxf + i[m])
        if isA:
            lam = np.stack([1 - ss, ss - tt, tt], 1)
            vs = ((0, 0), (2, 0), (2, 2))
        else:
            lam = np.stack([1 - tt, tt - ss, ss], 1)
            vs = ((0, 0), (0, 2), (2, 2))
        def node(di, dj):
            return (2 * ccy + dj) * Lxc + (2 * ccx + di)
        for a in range(3):
            rows.append(fine); cols.append(node(*vs[a])); vals.append(lam[:, a] * (2 * lam[:, a] - 1))
        for a, b in ((0, 1), (1, 2), (0, 2)):
            di, dj = (vs[a][0] + vs[b][0]) // 2, (vs[a][1] + vs[b][1]) // 2
            rows.append(fine); cols.append(node(di, dj)); vals.append(4 * lam[:, a] * lam[:, b])
    rows, cols, vals = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
    keep = np.abs(vals) > 1e-14
    P = sp.csr_matrix((vals[keep], (rows[keep], cols[keep])), shape=(Lxf * Lyf, Lxc * (2 * nyc + 1)))
    return sp.kron(P, sp.identity(2), format="csr")


class Hierarchy:
    def __init__(self, mesh, K, fixed, min_cells=2):
        self.A, self.P, self.free = [], [], []
        free = (~fixed).astype(float)
        F = sp.diags(free)
        A = (F @ K @ F + sp.diags(1.0 - free)).tocsr()
        self.A.append(A)
        self.free.append(free)
        nx, ny = mesh.nx, mesh.ny
        self.cells = [(nx, ny)]
        sides_fixed = fixed
        while nx % 2 == 0 and ny % 2 == 0 and min(nx, ny) // 2 >= 1 and max(nx, ny) > min_cells:
            nxc, nyc = nx // 2, ny // 2
            P = prolongation(nxc, nyc)
            mc = StructuredMesh(mesh.W, mesh.H, nxc, nyc)
            fc = mc.dirichlet_mask(self.fixed_sides)
            freec = (~fc).astype(float)
            P = (sp.diags(self.free[-1]) @ P @ sp.diags(freec)).tocsr()
            Ac = (P.T @ self.A[-1] @ P + sp.diags(1.0 - freec)).tocsr()
            self.P.append(P); self.A.append(Ac); self.free.append(freec)
            nx, ny = nxc, nyc
            self.cells.append((nx, ny))
        self.coarse = spla.splu(self.A[-1].tocsc())
        self.dinv = [1.0 / A.diagonal() for A in self.A]
        self.binv = [block_inverse(A) for A in self.A]
        self.lmax = [power_lmax(A, d) for A, d in zip(self.A, self.dinv)]
        self.lmax_b = [power_lmax(A, None, B) for A, B in zip(self.A, self.binv)]


def block_inverse(A):
    n = A.shape[0] // 2
    d = A.diagonal()
    a, c = d[0::2], d[1::2]
    idx = np.arange(n)
    b = np.asarray(A[2 * idx, 2 * idx + 1]).ravel()
    det = a * c - b * b
    rows = np.concatenate([2 * idx, 2 * idx, 2 * idx + 1, 2 * idx + 1])
    cols = np.concatenate([2 * idx, 2 * idx + 1, 2 * idx, 2 * idx + 1])
    vals = np.concatenate([c / det, -b / det, -b / det, a / det])
    return sp.csr_matrix((vals, (rows, cols)), shape=A.shape)


def power_lmax(A, dinv, B=None, its=60):
    rng = np.random.default_rng(0)
    v = rng.standard_normal(A.shape[0])
    lam = 1.0
    for _ in range(its):
        w = A @ v
        w = dinv * w if B is None else B @ w
        lam = np.linalg.norm(w) / np.linalg.norm(v)
        v = w / np.linalg.norm(w)
    return lam


OPT4 = {  # beta coefficients of the optimised fourth-kind smoother (Lottes 2022, table 1)
    1: [1.12500000000000],
    2: [1.02387287570313, 1.26408905371085],
    3: [1.00842544782028, 1.08867839208730, 1.33753125909618],
    4: [1.00391310427285, 1.04035811188593, 1.14863498546254, 1.38268869241000],
}


def smooth(A, apply_B, lmax, b, x, degree, kind, ratio=30.0, safety=1.1):
    """degree steps of a polynomial smoother; x None = zero initial guess."""
    if kind == "cheb1":
        hi = safety * lmax
        lo = hi / ratio
        theta, delta = 0.5 * (hi + lo), 0.5 * (hi - lo)
        sigma = theta / delta
        rho = 1.0 / sigma
        r = b.copy() if x is None else b - A @ x
        d = apply_B(r) / theta
        x = d.copy() if x is None else x + d
        for _ in range(1, degree):
            rho_new = 1.0 / (2 * sigma - rho)
            r = b - A @ x
            d = rho_new * rho * d + (2 * rho_new / delta) * apply_B(r)
            x = x + d
            rho = rho_new
        return x
    lam = safety * lmax
    beta = OPT4[degree] if kind == "opt4" else [1.0] * degree
    r = b.copy() if x is None else b - A @ x
    z = (4.0 / (3.0 * lam)) * apply_B(r)
    x = beta[0] * z if x is None else x + beta[0] * z
    for k in range(1, degree):
        r = r - A @ z  # residual of the UNWEIGHTED fourth-kind iterate, also for the optimised weights
        z = (2 * k - 1) / (2 * k + 3) * z + (8 * k + 4) / (2 * k + 3) / lam * apply_B(r)
        x = x + beta[k] * z
    return x


def make_vcycle(h, kind="cheb1", fine_degree=1, coarse_degree=3, block=False, ratio=30.0, safety=1.1, gamma=1,
                gamma_from=2, fine_kind=None, gamma_levels=None, count=None, window=None, light=0):
    nl = len(h.A)

    def B(l):
        if block:
            return lambda r: h.binv[l] @ r
        return lambda r: h.dinv[l] * r

    lm = h.lmax_b if block else h.lmax

    def cyc(l, b):
        if l == nl - 1:
            return h.coarse.solve(b)
        deg = fine_degree if l == 0 else (min(2, coarse_degree) if l <= light else coarse_degree)
        knd = (fine_kind or kind) if l == 0 else kind
        x = smooth(h.A[l], B(l), lm[l], b, None, deg, knd, ratio, safety)
        reps = gamma if l + 1 >= gamma_from and l + 1 < nl - 1 else 1
        if gamma_levels is not None:
            reps = gamma if (l + 1) in gamma_levels and l + 1 < nl - 1 else 1
        if window is not None:  # the library's automatic rule: cells on the short side of level l + 1
            m = min(h.cells[l + 1])
            reps = 2 if (window[0] <= m <= window[1] and l + 1 < nl - 1) else 1
        if count is not None:
            count[l] = count.get(l, 0) + 1
        for _ in range(reps):
            r = b - h.A[l] @ x
            x = x + h.P[l] @ cyc(l + 1, h.P[l].T @ r)
        return smooth(h.A[l], B(l), lm[l], b, x, deg, knd, ratio, safety)

    return lambda r: cyc(0, r)


def pcg(A, b, M, rtol=1e-10, maxit=400, x0=None):
    x = np.zeros_like(b) if x0 is None else x0.copy()
    r = b - A @ x
    z = M(r)
    p = z.copy()
    rz = r @ z
    bb = np.linalg.norm(b)
    for k in range(maxit):
        Ap = A @ p
        a = rz / (p @ Ap)
        x += a * p
        r -= a * Ap
        if np.linalg.norm(r) <= rtol * bb:
            return x, k + 1
        z = M(r)
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, maxit


VARIANTS3 = {  # the library's automatic choices (tm_engine.cu: repeats(), level_degree())
    "V-cycle, 1/3 steps (round-2 start)": dict(),
    "window 8..16 cells": dict(window=(8, 16)),
    "window 8..16, levels 1-2 two steps": dict(window=(8, 16), light=2),
    "window 4..32 cells": dict(window=(4, 32)),
    "window 4..32, levels 1-2 two steps": dict(window=(4, 32), light=2),
    "W on every level": dict(gamma=2, gamma_from=1),
}
VARIANTS2 = {
    "library: cheb1 1/3 V": dict(),
    "W all levels >= 2": dict(gamma=2, gamma_from=2),
    "repeat {2}": dict(gamma=2, gamma_levels={2}),
    "repeat {3}": dict(gamma=2, gamma_levels={3}),
    "repeat {4}": dict(gamma=2, gamma_levels={4}),
    "repeat {2,3}": dict(gamma=2, gamma_levels={2, 3}),
    "repeat {3,4}": dict(gamma=2, gamma_levels={3, 4}),
    "repeat {2,3,4}": dict(gamma=2, gamma_levels={2, 3, 4}),
    "repeat {2,4}": dict(gamma=2, gamma_levels={2, 4}),
    "repeat {1}": dict(gamma=2, gamma_levels={1}),
    "repeat {2,3,4,5}": dict(gamma=2, gamma_levels={2, 3, 4, 5}),
    "repeat {3,4,5}": dict(gamma=2, gamma_levels={3, 4, 5}),
    "repeat {4,5,6}": dict(gamma=2, gamma_levels={4, 5, 6}),
    "repeat {5,6,7}": dict(gamma=2, gamma_levels={5, 6, 7}),
    "repeat {1,2,3}": dict(gamma=2, gamma_levels={1, 2, 3}),
    "3x {2}": dict(gamma=3, gamma_levels={2}),
    "3x {3}": dict(gamma=3, gamma_levels={3}),
    "repeat {2}, coarse degree 2": dict(gamma=2, gamma_levels={2}, coarse_degree=2),
    "W all >= 2, coarse degree 2": dict(gamma=2, gamma_from=2, coarse_degree=2),
    "W all >= 2, coarse degree 1": dict(gamma=2, gamma_from=2, coarse_degree=1),
    "W all >= 1, coarse degree 2": dict(gamma=2, gamma_from=1, coarse_degree=2),
}


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    its = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    design = sys.argv[3] if len(sys.argv) > 3 else "bridge"
    cache = f"/tmp/tm_smoother_study_{design}_{N}_{its}.npz"
    s = OracleSolver(N * 2 if design == "bridge" else N, os.path.join(ROOT, "designs", f"{design}.json"))
    mesh = s.mesh
    print(f"{design}: {mesh.nx} x {mesh.ny} cells, {mesh.nu} dofs", flush=True)
    if os.path.exists(cache):
        z = np.load(cache)
        xis, us = z["xis"], z["us"]
    else:
        t0 = time.time()
        s.problem.set_penalization(3.0)
        psi = logit(s.rho)
        s.problem.calculate_objective(s.rho)
        xis, us = [s.problem.filtered_rho.copy()], [s.problem.u.copy()]
        for k in range(its):
            psi = s.step(psi.copy(), s.step_size_at_iter(k))
            s.rho = expit(psi)
            s.problem.calculate_objective(s.rho)
            xis.append(s.problem.filtered_rho.copy())
            us.append(s.problem.u.copy())
        xis, us = np.array(xis), np.array(us)
        np.savez(cache, xis=xis, us=us)
        print(f"oracle run {time.time() - t0:.0f} s", flush=True)
    pr = s.problem
    up = int(os.environ.get("STUDY_UPSAMPLE", "1"))
    if up > 1:  # the same designs, interpolated to an up-times finer mesh (more multigrid levels)
        from scipy.ndimage import zoom
        coarse = mesh
        s = OracleSolver(N * up * (2 if design == "bridge" else 1), os.path.join(ROOT, "designs", f"{design}.json"))
        mesh, pr = s.mesh, s.problem
        print(f"upsampled x{up}: {mesh.nx} x {mesh.ny} cells, {mesh.nu} dofs", flush=True)

        def up_p1(v):
            a = v.reshape(coarse.ny + 1, coarse.nx + 1)
            jj = np.linspace(0, coarse.ny, mesh.ny + 1)
            ii = np.linspace(0, coarse.nx, mesh.nx + 1)
            from scipy.interpolate import RegularGridInterpolator
            f = RegularGridInterpolator((np.arange(coarse.ny + 1), np.arange(coarse.nx + 1)), a)
            J, I = np.meshgrid(jj, ii, indexing="ij")
            return f(np.stack([J.ravel(), I.ravel()], 1))

        xis = np.array([up_p1(x) for x in xis[[0, its // 2, its]]])
        sel = {0: 0, its // 2: 1, its: 2}
        us = None
    Hierarchy.fixed_sides = s.design["fixed_sides"]
    variants = VARIANTS3 if os.environ.get("STUDY_SET") == "3" else VARIANTS2 if os.environ.get("STUDY_SET") == "2" else {
        "library: cheb1 1/3": dict(),
        "cheb1 1/3 ratio 10": dict(ratio=10.0),
        "cheb1 1/3 ratio 100": dict(ratio=100.0),
        "cheb4 fine cheb1(1), coarse 4th(3)": dict(kind="cheb4", fine_kind="cheb1"),
        "opt4 fine cheb1(1), coarse opt4(3)": dict(kind="opt4", fine_kind="cheb1"),
        "opt4 1/3 (fine opt4 too)": dict(kind="opt4"),
        "cheb4 1/3": dict(kind="cheb4"),
        "opt4 2/3": dict(kind="opt4", fine_degree=2),
        "cheb1 2/3": dict(fine_degree=2),
        "opt4 1/2": dict(kind="opt4", coarse_degree=2),
        "cheb1 1/2": dict(coarse_degree=2),
        "opt4 1/4": dict(kind="opt4", coarse_degree=4),
        "block cheb1 1/3": dict(block=True),
        "block opt4 1/3": dict(kind="opt4", block=True),
        "cheb1 1/3 W from level 2": dict(gamma=2, gamma_from=2),
        "opt4 1/3 W from level 2": dict(kind="opt4", gamma=2, gamma_from=2),
        "cheb1 1/3 W from level 1": dict(gamma=2, gamma_from=1),
    }
    for it in sorted({0, its // 2, its}):
        xi = xis[it] if up == 1 else xis[sel[it]]
        K = mesh.elasticity_matrix(xi, pr.lda, pr.mu, 3.0, 1e-6)
        h = Hierarchy(mesh, K, pr.fixed)
        b = pr.b_bc
        x0 = us[it - 1] if (it > 0 and us is not None) else None
        if it == 0 and up > 1:
            continue
        print(f"--- design after {it} iterations: xi in [{xi.min():.3f}, {xi.max():.3f}], levels {len(h.A)}, "
              f"lmax {['%.2f' % v for v in h.lmax]}", flush=True)
        for name, kw in variants.items():
            M = make_vcycle(h, **kw)
            _, n0 = pcg(h.A[0], b, M)
            n1 = pcg(h.A[0], b, M, x0=x0)[1] if x0 is not None else n0
            print(f"{name:44s} zero guess {n0:4d}   warm start {n1:4d}", flush=True)


if __name__ == "__main__":
    main()
