"""Element arithmetic of the CUDA kernels (tm_element.cuh, compiled for the host with g++)
against the quadrature-based oracle.  CPU only; guards the closed-form moment/strain algebra."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle.fem_oracle import StructuredMesh, triangle_rule

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hc():
    build = os.path.join(HERE, "_build")
    os.makedirs(build, exist_ok=True)
    so = os.path.join(build, "libelem_host.so")
    src = os.path.join(HERE, "hostcheck", "elem_host.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", src, "-o", so])
    lib = ctypes.CDLL(so)
    return lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


D = ctypes.c_double


@pytest.mark.parametrize("hx,hy", [(1.0, 1.0), (0.25, 0.25), (0.5, 0.2)])
def test_cell_matrix_matches_oracle(hc, hx, hy):
    rng = np.random.default_rng(3)
    xi4 = rng.random(4)
    lam, mu, m = 1.7, 0.9, 1e-6
    K = np.zeros((18, 18))
    hc.hc_cell_matrix(_ptr(xi4), D(m), D(lam), D(mu), D(hx), D(hy), _ptr(K))
    mesh = StructuredMesh(hx, hy, 1, 1)
    Ko = mesh.elasticity_matrix(xi4, lam, mu, 3.0, m).toarray()
    assert np.abs(K - K.T).max() < 1e-14
    assert np.abs(K - Ko).max() < 1e-13 * np.abs(Ko).max()


def test_moments_match_quadrature(hc):
    rng = np.random.default_rng(5)
    pts, wts = triangle_rule(5)
    for _ in range(5):
        xi = rng.random(3)
        m = 1e-6
        w = np.zeros(6)
        hc.hc_moments(_ptr(xi), D(m), _ptr(w))
        r = m + (1 - m) * (pts @ xi) ** 3
        pairs = [(0, 0), (1, 1), (2, 2), (0, 1), (1, 2), (0, 2)]
        ref = np.array([0.5 * np.sum(wts * r * pts[:, a] * pts[:, b]) for a, b in pairs])
        assert np.abs(w - ref).max() < 1e-15


def test_cell_sensitivity_matches_oracle(hc):
    rng = np.random.default_rng(7)
    hx, hy = 0.5, 0.25
    lam, mu, m = 2.0, 1.5, 1e-6
    xi4 = rng.random(4)
    u = rng.standard_normal(18)
    g4 = np.zeros(4)
    hc.hc_cell_sensitivity(_ptr(u), _ptr(xi4), D(m), D(lam), D(mu), D(hx), D(hy), _ptr(g4))
    mesh = StructuredMesh(hx, hy, 1, 1)
    ref = mesh.sensitivity_rhs(u, xi4, lam, mu, 3.0, m, nq=5)
    assert np.abs(g4 - ref).max() < 1e-12 * np.abs(ref).max()


# ---- general SIMP exponent: integer p exact, other p on the oracle's 16-point rule
@pytest.mark.parametrize("p", [1.0, 2.0, 3.0, 4.0, 5.0, 8.0, 16.0])
def test_general_moments_integer_exponent_exact(hc, p):
    rng = np.random.default_rng(int(p))
    pts, wts = triangle_rule(12)  # exact to degree 22 >= p + 2
    pairs = [(0, 0), (1, 1), (2, 2), (0, 1), (1, 2), (0, 2)]
    for _ in range(4):
        xi, m, w = rng.random(3), 1e-6, np.zeros(6)
        hc.hc_moments_general(_ptr(xi), D(m), D(p), _ptr(w))
        r = m + (1 - m) * (pts @ xi) ** p
        ref = np.array([0.5 * np.sum(wts * r * pts[:, a] * pts[:, b]) for a, b in pairs])
        assert np.abs(w - ref).max() < 2e-15
        if p == 3.0:
            w3 = np.zeros(6)
            hc.hc_moments(_ptr(xi), D(m), _ptr(w3))
            assert np.abs(w - w3).max() < 1e-16


@pytest.mark.parametrize("p", [0.5, 1.5, 2.5, 3.7, 17.0])
def test_general_moments_other_exponents_use_the_oracle_rule(hc, p):
    rng = np.random.default_rng(11)
    pts, wts = triangle_rule(4)
    pairs = [(0, 0), (1, 1), (2, 2), (0, 1), (1, 2), (0, 2)]
    xi, m, w = rng.random(3), 1e-6, np.zeros(6)
    hc.hc_moments_general(_ptr(xi), D(m), D(p), _ptr(w))
    r = m + (1 - m) * (pts @ xi) ** p
    ref = np.array([0.5 * np.sum(wts * r * pts[:, a] * pts[:, b]) for a, b in pairs])
    assert np.abs(w - ref).max() < 1e-15


@pytest.mark.parametrize("p,nq", [(1.0, 6), (2.0, 6), (4.0, 6), (6.0, 8), (2.5, 4), (0.75, 4)])
def test_general_cell_matrix_and_sensitivity_match_oracle(hc, p, nq):
    rng = np.random.default_rng(13)
    hx, hy, lam, mu, m = 0.5, 0.25, 2.0, 1.5, 1e-6
    xi4 = 0.05 + 0.95 * rng.random(4)
    mesh = StructuredMesh(hx, hy, 1, 1)
    K = np.zeros((18, 18))
    hc.hc_cell_matrix_general(_ptr(xi4), D(m), D(p), D(lam), D(mu), D(hx), D(hy), _ptr(K))
    Ko = mesh.elasticity_matrix(xi4, lam, mu, p, m, nq=nq).toarray()
    assert np.abs(K - Ko).max() < 1e-13 * np.abs(Ko).max()
    u, g4 = rng.standard_normal(18), np.zeros(4)
    hc.hc_cell_sensitivity_general(_ptr(u), _ptr(xi4), D(m), D(p), D(lam), D(mu), D(hx), D(hy), _ptr(g4))
    ref = mesh.sensitivity_rhs(u, xi4, lam, mu, p, m, nq=nq)
    assert np.abs(g4 - ref).max() < 1e-12 * np.abs(ref).max()
