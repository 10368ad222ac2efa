"""CPU oracle for the FEM elasticity inner loop of Emilinya/topomax.

TEST INFRASTRUCTURE ONLY.  Nothing under ``topomax_b200/`` imports this module;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may.  It is a plain numpy/scipy fp64 restatement of what the
reference computes through legacy FEniCS (``dolfin`` 2019.1.x + FFC + PETSc/MUMPS,
un-vendored and absent from this image, so the reference itself cannot be run):

* mesh / spaces ........ FEM_src/solver.py:38-49, FEM_src/elasisity_problem.py:194-196
  (``RectangleMesh`` default "right" diagonal, P1 control space, vector-P2 state space)
* Helmholtz filter ..... FEM_src/filter.py:27-41 + FEM_src/pde_solver.py:106-133
* state operator ....... FEM_src/elasisity_problem.py:112-118, src/penalizers.py:29-46
* load vector .......... FEM_src/elasisity_problem.py:20-73,120-124
* Dirichlet rows ....... FEM_src/elasisity_problem.py:171-192, FEM_src/domains.py:18-38
* compliance ........... FEM_src/elasisity_problem.py:152-166
* filtered sensitivity . FEM_src/elasisity_problem.py:134-150
* nodal integral ....... FEM_src/solver.py:81-84

Parity status: PINNED for the body-force path by the reference's own golden fixture
``tests/test_data/FEM/triangle/data/correct_{data,rho}.dat`` (tests/test_elasticity_solver.py:30-55)
and by ``tests/test_filter.py:25-60`` for the filter; the traction-load path, displacement
fields and p != 3 are PARITY UNPINNED (no reference fixture exists).

Everything is assembled element by element with numerical quadrature on the actual
triangle coordinates (no structured-mesh shortcuts), then solved with a sparse direct
solver (SuperLU playing MUMPS' role).

Vector layouts (shared with the CUDA library so arrays are directly comparable):
* P1 fields: ``v = iy*(nx+1) + ix``                      (vertex grid, row-major)
* P2 fields: node ``n = j*(2nx+1) + i`` on the half-step lattice, dof ``2n + component``.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

DOLFIN_EPS = 3.0e-16  # dolfin/common/constants.h, used by df.between / df.near


# --------------------------------------------------------------------------------------
# quadrature
# --------------------------------------------------------------------------------------
def triangle_rule(n: int = 4):
    """Collapsed (Duffy) Gauss rule on the unit triangle, exact to degree 2n-2.

    Returns barycentric points ``(nq,3)`` and weights summing to 1 (area-normalised).
    n=4 is exact for degree 6 >= the degree-5 integrands of the p=3 path.
    """
    g, w = np.polynomial.legendre.leggauss(n)
    g = 0.5 * (g + 1.0)
    w = 0.5 * w
    pts, wts = [], []
    for a, wa in zip(g, w):
        for b, wb in zip(g, w):
            # (x, y) = (a, b (1-a)),  jacobian (1-a), reference area 1/2
            x, y = a, b * (1.0 - a)
            pts.append((1.0 - x - y, x, y))
            wts.append(wa * wb * (1.0 - a) * 2.0)
    return np.array(pts), np.array(wts)


def segment_rule(n: int = 3):
    g, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (g + 1.0), 0.5 * w


# P2 basis on barycentrics; local nodes: 0,1,2 vertices; 3=mid(0,1); 4=mid(1,2); 5=mid(0,2)
_P2_EDGES = ((0, 1), (1, 2), (0, 2))


def p2_basis(lam):
    l0, l1, l2 = lam
    return np.array([
        l0 * (2 * l0 - 1), l1 * (2 * l1 - 1), l2 * (2 * l2 - 1),
        4 * l0 * l1, 4 * l1 * l2, 4 * l0 * l2,
    ])


def p2_basis_grad(lam, grad_lam):
    """grad phi_k at barycentric point ``lam``; ``grad_lam`` is (3,2). Returns (6,2)."""
    out = np.zeros((6, 2))
    for a in range(3):
        out[a] = (4 * lam[a] - 1) * grad_lam[a]
    for k, (a, b) in enumerate(_P2_EDGES):
        out[3 + k] = 4 * (lam[a] * grad_lam[b] + lam[b] * grad_lam[a])
    return out


class StructuredMesh:
    """``df.RectangleMesh(Point(0,0), Point(W,H), nx, ny)`` with the default "right" diagonal.

    Vertex coordinates follow dolfin's generator: ``x = (ix*W)/nx`` (RectangleMesh.cpp
    build_tri), midpoints are the mean of their two end vertices (affine map of the P2
    reference dof points).  Each cell splits along the diagonal v0->v3 into
    T_A=(v0,v1,v3) and T_B=(v0,v2,v3)  (SURVEY App. A.1).
    """

    def __init__(self, W: float, H: float, nx: int, ny: int):
        self.W, self.H, self.nx, self.ny = float(W), float(H), int(nx), int(ny)
        self.n1 = (nx + 1) * (ny + 1)
        self.Lx, self.Ly = 2 * nx + 1, 2 * ny + 1
        self.n2 = self.Lx * self.Ly
        self.nu = 2 * self.n2

        xv = (np.arange(nx + 1, dtype=np.float64) * self.W) / nx
        yv = (np.arange(ny + 1, dtype=np.float64) * self.H) / ny
        self.xv, self.yv = xv, yv
        # lattice coordinates: even index = vertex, odd = mean of neighbours
        xl = np.empty(self.Lx)
        xl[0::2] = xv
        xl[1::2] = 0.5 * xv[:-1] + 0.5 * xv[1:]
        yl = np.empty(self.Ly)
        yl[0::2] = yv
        yl[1::2] = 0.5 * yv[:-1] + 0.5 * yv[1:]
        self.xl, self.yl = xl, yl

        ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
        ix, iy = ix.ravel(), iy.ravel()
        v0 = iy * (nx + 1) + ix
        v1, v2, v3 = v0 + 1, v0 + (nx + 1), v0 + (nx + 1) + 1
        # P1 connectivity per orientation
        self.tri_v = {"A": np.stack([v0, v1, v3], 1), "B": np.stack([v0, v2, v3], 1)}
        # lattice coordinates (i,j) of the three vertices of each triangle
        lat = {
            "A": ((0, 0), (2, 0), (2, 2)),
            "B": ((0, 0), (0, 2), (2, 2)),
        }
        self.tri_n = {}
        for t, vs in lat.items():
            cols = []
            for (di, dj) in vs:
                cols.append((2 * iy + dj) * self.Lx + (2 * ix + di))
            for (a, b) in _P2_EDGES:
                di = (vs[a][0] + vs[b][0]) // 2
                dj = (vs[a][1] + vs[b][1]) // 2
                cols.append((2 * iy + dj) * self.Lx + (2 * ix + di))
            self.tri_n[t] = np.stack(cols, 1)
        # reference triangle geometry per orientation (all triangles of a type are translates)
        self.geom = {}
        for t, vs in lat.items():
            P = np.array([[xl[di], yl[dj]] for (di, dj) in vs])  # cell (0,0)
            J = np.array([P[1] - P[0], P[2] - P[0]]).T
            area = 0.5 * abs(np.linalg.det(J))
            Jinv = np.linalg.inv(J)
            # lam1, lam2 are the reference coords; lam0 = 1 - lam1 - lam2
            g12 = Jinv  # rows: grad lam1, grad lam2
            grad_lam = np.array([-(g12[0] + g12[1]), g12[0], g12[1]])
            self.geom[t] = (area, grad_lam)

    # ---------------------------------------------------------------- P1 operators
    def p1_matrices(self):
        """Consistent P1 stiffness K1 and mass M1 (exact), natural BC. FEM_src/filter.py:27-33."""
        rows, cols, kv, mv = [], [], [], []
        for t in ("A", "B"):
            area, gl = self.geom[t]
            Ke = area * gl @ gl.T
            Me = area / 12.0 * (np.ones((3, 3)) + np.eye(3))
            conn = self.tri_v[t]
            r = np.repeat(conn, 3, axis=1)
            c = np.tile(conn, (1, 3))
            rows.append(r.ravel()); cols.append(c.ravel())
            kv.append(np.tile(Ke.ravel(), conn.shape[0]))
            mv.append(np.tile(Me.ravel(), conn.shape[0]))
        rows, cols = np.concatenate(rows), np.concatenate(cols)
        K1 = sp.csr_matrix((np.concatenate(kv), (rows, cols)), shape=(self.n1, self.n1))
        M1 = sp.csr_matrix((np.concatenate(mv), (rows, cols)), shape=(self.n1, self.n1))
        return K1, M1

    def nodal_weights(self):
        """w = M1 . 1 : ``integrate(values) = w . values`` (FEM_src/solver.py:81-84)."""
        _, M1 = self.p1_matrices()
        return np.asarray(M1.sum(axis=1)).ravel()

    # ---------------------------------------------------------------- P2 helpers
    def _strain_rows(self, t, lam):
        """3x12 matrix mapping local dofs (node-major, 2 comps) to (e_xx, e_yy, gamma_xy)."""
        _, gl = self.geom[t]
        g = p2_basis_grad(lam, gl)  # (6,2)
        B = np.zeros((3, 12))
        B[0, 0::2] = g[:, 0]
        B[1, 1::2] = g[:, 1]
        B[2, 0::2] = g[:, 1]
        B[2, 1::2] = g[:, 0]
        return B

    def _local_dofs(self, t):
        n = self.tri_n[t]
        d = np.empty((n.shape[0], 12), dtype=np.int64)
        d[:, 0::2] = 2 * n
        d[:, 1::2] = 2 * n + 1
        return d

    def elasticity_matrix(self, xi, lam_, mu, p=3.0, m=1e-6, nq=4):
        """K(xi)_ij = int r(xi_h) [lam div phi_j div phi_i + 2 mu eps(phi_j):eps(phi_i)].

        FEM_src/elasisity_problem.py:112-118 with r = m + (1-m) xi^p (src/penalizers.py:36-40).
        No boundary conditions applied.
        """
        D = np.array([[lam_ + 2 * mu, lam_, 0.0], [lam_, lam_ + 2 * mu, 0.0], [0.0, 0.0, mu]])
        pts, wts = triangle_rule(nq)
        rows, cols, vals = [], [], []
        for t in ("A", "B"):
            area, _ = self.geom[t]
            G = np.stack([self._strain_rows(t, q).T @ D @ self._strain_rows(t, q) for q in pts])  # (nq,12,12)
            xv = xi[self.tri_v[t]]  # (nt,3)
            xq = xv @ pts.T  # (nt,nq)
            cq = (m + (1.0 - m) * xq**p) * (wts * area)  # (nt,nq)
            Ke = cq @ G.reshape(len(pts), 144)  # (nt,144)
            d = self._local_dofs(t)
            rows.append(np.repeat(d, 12, axis=1).ravel())
            cols.append(np.tile(d, (1, 12)).ravel())
            vals.append(Ke.ravel())
        return sp.csr_matrix(
            (np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
            shape=(self.nu, self.nu),
        )

    def p2_mass(self, nq=4):
        """Scalar P2 consistent mass matrix on the lattice (n2 x n2)."""
        pts, wts = triangle_rule(nq)
        rows, cols, vals = [], [], []
        for t in ("A", "B"):
            area, _ = self.geom[t]
            Me = sum(w * np.outer(p2_basis(q), p2_basis(q)) for q, w in zip(pts, wts)) * area
            n = self.tri_n[t]
            rows.append(np.repeat(n, 6, axis=1).ravel())
            cols.append(np.tile(n, (1, 6)).ravel())
            vals.append(np.tile(Me.ravel(), n.shape[0]))
        return sp.csr_matrix(
            (np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
            shape=(self.n2, self.n2),
        )

    def node_coordinates(self):
        """(n2,) x and y of every lattice node, dolfin-style floating point."""
        X, Y = np.meshgrid(self.xl, self.yl, indexing="xy")
        return X.ravel(), Y.ravel()

    # ---------------------------------------------------------------- loads
    def load_vector(self, body_force=None, tractions=None):
        """b = int f_h.v dx + int t_h.v ds with f_h, t_h the P2 nodal interpolants of the
        reference's indicator ``UserExpression``s (degree-2 default element).

        ``body_force``: (cx, cy, radius, fx, fy) or None   (FEM_src/elasisity_problem.py:26-33)
        ``tractions``: list of (side, center, length, tx, ty), side in Left/Right/Top/Bottom
                       (FEM_src/elasisity_problem.py:47-70)
        """
        b = np.zeros(self.nu)
        X, Y = self.node_coordinates()
        if body_force is not None:
            cx, cy, rad, fx, fy = body_force
            dist = np.sqrt((X - cx) ** 2 + (Y - cy) ** 2)
            inside = (dist < rad).astype(np.float64)
            M2 = self.p2_mass()
            mi = M2 @ inside
            b[0::2] += fx * mi
            b[1::2] += fy * mi
        tractions = list(tractions or [])
        for (side, *_rest) in tractions:
            if side not in ("Left", "Right", "Top", "Bottom"):
                raise ValueError(f"Malformed side: {side}")

        def nodal_value(x, y):
            """TractionExpression.eval at one node: every traction whose side passes through the node and
            whose window contains it -- at a corner that includes tractions of BOTH adjacent sides."""
            vx = vy = 0.0
            for (side, center, length, tx, ty) in tractions:
                lo, hi = center - length / 2, center + length / 2
                if side == "Left":
                    hit = x == 0.0 and lo - DOLFIN_EPS <= y <= hi + DOLFIN_EPS
                elif side == "Right":
                    hit = x == self.W and lo - DOLFIN_EPS <= y <= hi + DOLFIN_EPS
                elif side == "Top":
                    hit = y == self.H and lo - DOLFIN_EPS <= x <= hi + DOLFIN_EPS
                else:
                    hit = y == 0.0 and lo - DOLFIN_EPS <= x <= hi + DOLFIN_EPS
                if hit:
                    vx += tx
                    vy += ty
            return vx, vy

        if tractions:
            # ds runs over all four sides; on each boundary edge the P2 interpolant of the expression is
            # fixed by its values at the edge's three nodes: exact 1-D P2 mass (v0, mid, v1), 3-point Gauss
            g, w = segment_rule(3)
            phi = np.array([[(1 - x) * (1 - 2 * x), 4 * x * (1 - x), x * (2 * x - 1)] for x in g])
            for side in ("Left", "Right", "Bottom", "Top"):
                if side in ("Left", "Right"):
                    i = 0 if side == "Left" else self.Lx - 1
                    s = self.yl
                    nodes = np.arange(self.Ly) * self.Lx + i
                    vals = np.array([nodal_value(self.xl[i], y) for y in s])
                else:
                    j = 0 if side == "Bottom" else self.Ly - 1
                    s = self.xl
                    nodes = j * self.Lx + np.arange(self.Lx)
                    vals = np.array([nodal_value(x, self.yl[j]) for x in s])
                if not vals.any():
                    continue
                acc = np.zeros_like(vals)
                for e in range((len(s) - 1) // 2):
                    loc = [2 * e, 2 * e + 1, 2 * e + 2]
                    if not vals[loc].any():
                        continue
                    hlen = s[2 * e + 2] - s[2 * e]
                    Me = hlen * (phi.T * w) @ phi
                    acc[loc] += Me @ vals[loc]
                b[2 * nodes] += acc[:, 0]
                b[2 * nodes + 1] += acc[:, 1]
        return b

    def dirichlet_mask(self, fixed_sides):
        """Boolean (nu,) mask of displacement dofs on the fixed sides."""
        mask = np.zeros((self.Ly, self.Lx), dtype=bool)
        for side in fixed_sides:
            if side == "Left":
                mask[:, 0] = True
            elif side == "Right":
                mask[:, -1] = True
            elif side == "Bottom":
                mask[0, :] = True
            elif side == "Top":
                mask[-1, :] = True
            else:
                raise ValueError(f"Malformed side: {side}")
        return np.repeat(mask.ravel(), 2)

    # ---------------------------------------------------------------- sensitivity
    def sensitivity_rhs(self, u, xi, lam_, mu, p=3.0, m=1e-6, nq=4):
        """b^g_i = int -r'(xi_h) (lam (div u)^2 + 2 mu eps(u):eps(u)) phi_i^{P1} dx.

        FEM_src/elasisity_problem.py:146-150, r' = p xi^(p-1) (1-m) (src/penalizers.py:42-46).
        """
        pts, wts = triangle_rule(nq)
        out = np.zeros(self.n1)
        for t in ("A", "B"):
            area, _ = self.geom[t]
            ue = u[self._local_dofs(t)]  # (nt,12)
            xv = xi[self.tri_v[t]]
            for q, w in zip(pts, wts):
                B = self._strain_rows(t, q)
                e = ue @ B.T  # (nt,3)
                energy = lam_ * (e[:, 0] + e[:, 1]) ** 2 + 2 * mu * (
                    e[:, 0] ** 2 + e[:, 1] ** 2 + 0.5 * e[:, 2] ** 2
                )
                xq = xv @ q
                g = -(p * xq ** (p - 1) * (1 - m)) * energy * (w * area)
                for c in range(3):
                    out += np.bincount(self.tri_v[t][:, c], weights=g * q[c], minlength=self.n1)
        return out


def lame(E, nu):
    """FEM_src/elasisity_problem.py:98-99."""
    mu = E / (2 * (1 + nu))
    lda = mu * nu / (0.5 - nu)
    return lda, mu


def nested_dissection_order(Lx, Ly, leaf=6):
    """Fill-reducing ordering of the Lx x Ly P2 lattice for the direct solver (MUMPS would get
    one from METIS/SCOTCH): recursive bisection with single even lattice lines as separators
    (nodes two lattice steps apart across an even line share no cell, hence no matrix entry)."""
    pieces = []

    def box(i0, i1, j0, j1):
        jj, ii = np.meshgrid(np.arange(j0, j1 + 1), np.arange(i0, i1 + 1), indexing="ij")
        pieces.append((jj * Lx + ii).ravel())

    def split(lo, hi):
        m = (lo + hi) // 2
        m -= m & 1
        if m <= lo:
            m = lo + (2 if (lo & 1) == 0 else 1)
        return m

    def rec(i0, i1, j0, j1):
        w, h = i1 - i0 + 1, j1 - j0 + 1
        if w <= 0 or h <= 0:
            return
        if w <= leaf and h <= leaf:
            return box(i0, i1, j0, j1)
        if w >= h:
            m = split(i0, i1)
            if m >= i1:
                return box(i0, i1, j0, j1)
            rec(i0, m - 1, j0, j1)
            rec(m + 1, i1, j0, j1)
            pieces.append(np.arange(j0, j1 + 1) * Lx + m)
        else:
            m = split(j0, j1)
            if m >= j1:
                return box(i0, i1, j0, j1)
            rec(i0, i1, j0, m - 1)
            rec(i0, i1, m + 1, j1)
            pieces.append(m * Lx + np.arange(i0, i1 + 1))

    rec(0, Lx - 1, 0, Ly - 1)
    order = np.concatenate(pieces)
    assert order.size == Lx * Ly
    return order


def solve_spd(A, b, free=None, lattice=None):
    """Sparse direct solve (SuperLU in MUMPS' role, FEM_src/pde_solver.py:130-131).

    With ``free`` given, Dirichlet dofs are eliminated symmetrically: identical to the
    reference's row-zero/unit-diagonal ``bc.apply`` because the prescribed value is 0.
    ``lattice=(Lx, Ly)`` (vector-P2 systems) switches on the nested-dissection ordering, which
    is what keeps the factorisation affordable beyond ~1e5 dofs.
    """
    if free is None:
        return spla.splu(A.tocsc()).solve(b)
    x = np.zeros_like(b)
    if lattice is not None and A.shape[0] > 20000:
        nodes = nested_dissection_order(*lattice)
        idx = np.stack([2 * nodes, 2 * nodes + 1], 1).ravel()
        idx = idx[free[idx]]
        Aff = A[idx][:, idx].tocsc()
        lu = spla.splu(Aff, permc_spec="NATURAL", diag_pivot_thresh=0.0,
                       options=dict(SymmetricMode=True))
    else:
        idx = np.flatnonzero(free)
        Aff = A[idx][:, idx].tocsc()
        lu = spla.splu(Aff, permc_spec="MMD_AT_PLUS_A")
    x[idx] = lu.solve(b[idx])
    return x


def l2_error_p1(mesh, xi, exact, nq=6):
    """|| xi_h - exact ||_L2 by quadrature (the role of df.errornorm in tests/test_filter.py)."""
    pts, wts = triangle_rule(nq)
    X, Y = np.meshgrid(mesh.xv, mesh.yv, indexing="xy")
    X, Y = X.ravel(), Y.ravel()
    err = 0.0
    for t in ("A", "B"):
        area, _ = mesh.geom[t]
        v = mesh.tri_v[t]
        for q, w in zip(pts, wts):
            xq, yq, fq = X[v] @ q, Y[v] @ q, xi[v] @ q
            err += w * area * np.sum((fq - exact(xq, yq)) ** 2)
    return np.sqrt(err)


# --------------------------------------------------------------------------------------
# point evaluation (the role of dolfin's ``f(x, y)`` in sample_function, FEM_src/utils.py:112-162)
# --------------------------------------------------------------------------------------
def evaluate_field(mesh: StructuredMesh, values, degree: int, xs, ys):
    """Values of a P1 (degree 1, returns (n,1)) or vector-P2 (degree 2, returns (n,2)) field at
    the points (xs[k], ys[k]).  Geometric, as dolfin does it: find a triangle that contains the
    point (barycentric coordinates from the triangle's actual vertex coordinates, all >= -1e-12),
    then sum basis function x dof.  Points on shared edges give the same value from either side
    because the spaces are continuous."""
    xs, ys = np.atleast_1d(np.asarray(xs, float)), np.atleast_1d(np.asarray(ys, float))
    values = np.asarray(values, float)
    X, Y = np.meshgrid(mesh.xv, mesh.yv, indexing="xy")
    X, Y = X.ravel(), Y.ravel()
    # candidate cell by bisection on the vertex coordinates (the bounding-box tree's job)
    cx = np.clip(np.searchsorted(mesh.xv, xs, side="right") - 1, 0, mesh.nx - 1)
    cy = np.clip(np.searchsorted(mesh.yv, ys, side="right") - 1, 0, mesh.ny - 1)
    cell = cy * mesh.nx + cx
    out = np.full((xs.size, degree), np.nan)
    done = np.zeros(xs.size, bool)
    for t in ("A", "B"):
        v = mesh.tri_v[t][cell]  # (n, 3) vertex ids
        p0 = np.stack([X[v[:, 0]], Y[v[:, 0]]], 1)
        e1 = np.stack([X[v[:, 1]], Y[v[:, 1]]], 1) - p0
        e2 = np.stack([X[v[:, 2]], Y[v[:, 2]]], 1) - p0
        d = np.stack([xs, ys], 1) - p0
        det = e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]
        l1 = (d[:, 0] * e2[:, 1] - d[:, 1] * e2[:, 0]) / det
        l2 = (e1[:, 0] * d[:, 1] - e1[:, 1] * d[:, 0]) / det
        lam = np.stack([1.0 - l1 - l2, l1, l2], 1)
        inside = np.all(lam >= -1e-12, axis=1) & ~done
        if degree == 1:
            out[inside, 0] = np.sum(lam[inside] * values[v[inside]], axis=1)
        else:
            nodes = mesh.tri_n[t][cell[inside]]  # (m, 6) lattice node ids
            phi = p2_basis(lam[inside].T).T  # (m, 6)
            for c in range(2):
                out[inside, c] = np.sum(phi * values[2 * nodes + c], axis=1)
        done |= inside
    if not done.all():
        raise ValueError("evaluate_field: a point lies outside the mesh")
    return out


def sample_function(mesh: StructuredMesh, values, degree: int, points: int, sample_type: str, N: int):
    """Restatement of FEM_src/utils.py:112-162 on the oracle mesh: returns (domain_rays,
    output_grid[nsy, nsx, degree])."""
    domain_size = (mesh.W, mesh.H)
    multiplier = int(np.ceil(points / N))
    domain_samples = [int(s * N * multiplier) for s in domain_size]
    if sample_type == "edges":
        domain_samples = [ns + 1 for ns in domain_samples]
    domain_rays = [np.linspace(0, s, ns) for s, ns in zip(domain_size, domain_samples)]
    xi, yi = np.meshgrid(np.arange(domain_samples[0]), np.arange(domain_samples[1]), indexing="xy")
    if sample_type == "center":
        x, y = (0.5 + xi) / (multiplier * N), (0.5 + yi) / (multiplier * N)
    elif sample_type == "edges":
        x, y = xi / (multiplier * N), yi / (multiplier * N)
    else:
        raise ValueError(f"Unknown sample_type: {sample_type}. sample_type must be either 'center' or 'edges'")
    grid = evaluate_field(mesh, values, degree, x.ravel(), y.ravel())
    return domain_rays, grid.reshape(domain_samples[1], domain_samples[0], degree)
