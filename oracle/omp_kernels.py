"""ctypes front of ``oracle/c/elast_omp.c``: the C + OpenMP CPU kernel baseline (matrix-free
penalised P2 elasticity operator + Jacobi-PCG on all host cores).

TEST / MEASUREMENT INFRASTRUCTURE ONLY (see ``oracle/fem_oracle.py`` header).  The shared
object is built into ``oracle/_build/`` (git-ignored, travels with gpurun snapshots).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from .fem_oracle import triangle_rule

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c", "elast_omp.c")
OUT = os.path.join(HERE, "_build", "liboracle_omp.so")
# no -march=native: the object is built in one container and may run on another host
CFLAGS = ["-O3", "-fopenmp", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off"]

SIDES = ("Left", "Right", "Bottom", "Top")


class _Problem(ctypes.Structure):
    _fields_ = [("nx", ctypes.c_int), ("ny", ctypes.c_int), ("W", ctypes.c_double), ("H", ctypes.c_double),
                ("lam", ctypes.c_double), ("mu", ctypes.c_double), ("p", ctypes.c_double), ("m", ctypes.c_double),
                ("fixed", ctypes.c_int * 4), ("nq", ctypes.c_int),
                ("pts", ctypes.POINTER(ctypes.c_double)), ("wts", ctypes.POINTER(ctypes.c_double))]


def build(force: bool = False) -> str:
    if force or not os.path.isfile(OUT) or os.path.getmtime(OUT) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.run(["gcc", *CFLAGS, SRC, "-o", OUT, "-lm"], check=True)
    return OUT


_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        dp = ctypes.POINTER(ctypes.c_double)
        pp = ctypes.POINTER(_Problem)
        lib.oc_num_threads.restype = ctypes.c_int
        lib.oc_set_num_threads.argtypes = [ctypes.c_int]
        lib.oc_elast_apply.argtypes = [pp, dp, dp, dp]
        lib.oc_elast_diag.argtypes = [pp, dp, dp]
        lib.oc_jacobi_pcg.argtypes = [pp, dp, dp, dp, ctypes.c_double, ctypes.c_int,
                                      ctypes.POINTER(ctypes.c_int), dp, dp]
        _lib = lib
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


class OmpElasticity:
    """The state operator of one mesh/material/penalty on the host cores."""

    def __init__(self, W, H, nx, ny, lam, mu, fixed_sides, p=3.0, m=1e-6, nq=4, threads=0):
        self.lib = _load()
        if threads:
            self.lib.oc_set_num_threads(int(threads))
        pts, wts = triangle_rule(nq)
        self._pts = np.ascontiguousarray(pts, dtype=np.float64)
        self._wts = np.ascontiguousarray(wts, dtype=np.float64)
        for s in fixed_sides:
            if s not in SIDES:
                raise ValueError(f"Malformed side: {s}")
        self.prob = _Problem(int(nx), int(ny), float(W), float(H), float(lam), float(mu), float(p), float(m),
                             (ctypes.c_int * 4)(*[int(s in fixed_sides) for s in SIDES]), len(self._wts),
                             _ptr(self._pts), _ptr(self._wts))
        self.n1 = (nx + 1) * (ny + 1)
        self.nu = 2 * (2 * nx + 1) * (2 * ny + 1)

    @property
    def threads(self) -> int:
        return int(self.lib.oc_num_threads())

    def _check(self, xi, *vecs):
        xi = np.ascontiguousarray(xi, dtype=np.float64)
        assert xi.size == self.n1
        out = [xi]
        for v in vecs:
            v = np.ascontiguousarray(v, dtype=np.float64)
            assert v.size == self.nu
            out.append(v)
        return out

    def apply(self, xi, x):
        xi, x = self._check(xi, x)
        y = np.empty(self.nu)
        rc = self.lib.oc_elast_apply(ctypes.byref(self.prob), _ptr(xi), _ptr(x), _ptr(y))
        if rc:
            raise RuntimeError(f"oc_elast_apply failed ({rc})")
        return y

    def diagonal(self, xi):
        (xi,) = self._check(xi)
        d = np.empty(self.nu)
        rc = self.lib.oc_elast_diag(ctypes.byref(self.prob), _ptr(xi), _ptr(d))
        if rc:
            raise RuntimeError(f"oc_elast_diag failed ({rc})")
        return d

    def jacobi_pcg(self, xi, b, rtol=1e-10, maxit=100000):
        """Returns (u, iterations, relative residual, seconds of the iteration loop)."""
        xi, b = self._check(xi, b)
        u = np.empty(self.nu)
        its, rel, sec = ctypes.c_int(0), ctypes.c_double(0.0), ctypes.c_double(0.0)
        rc = self.lib.oc_jacobi_pcg(ctypes.byref(self.prob), _ptr(xi), _ptr(b), _ptr(u), float(rtol), int(maxit),
                                    ctypes.byref(its), ctypes.byref(rel), ctypes.byref(sec))
        if rc:
            raise RuntimeError(f"oc_jacobi_pcg failed ({rc})")
        return u, its.value, rel.value, sec.value
