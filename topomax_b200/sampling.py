"""Point sampling of P1 / vector-P2 fields on a regular grid, on the device
(reference: ``sample_function``, FEM_src/utils.py:112-162, consumed by plot.py:54-69).

The reference evaluates ``f(x, y)`` point by point in a Python double loop through dolfin's
bounding-box tree; here one ``sample_field_kernel`` launch (``tm_sample_field``) finds the cell by
index arithmetic and evaluates the P1 / P2 basis, and the grid of values crosses PCIe once.
"""
from __future__ import annotations

import numpy as np

from .mesh import Function, RectangleMesh

_ENGINES: dict = {}


def mesh_to_N(mesh: RectangleMesh) -> int:
    """Elements per unit length, recovered from the cell diameter as the reference does
    (FEM_src/utils.py:29-37; hmin is the hypotenuse of a mesh square's triangle)."""
    n = 1 / (mesh.hmin() / np.sqrt(2))
    if abs(n - int(round(n))) / n > 1e-10:
        print(f"save_function: got non-integer N: {n}, this could result in wrong data getting saved")
    return int(round(n))


def mesh_to_domain_size(mesh: RectangleMesh):
    """(width, height) of the mesh (reference: FEM_src/utils.py:40-44)."""
    return mesh.domain_size


def _sampling_engine(f: Function):
    """A bare engine on the function's mesh (creating one allocates only scratch; no hierarchy
    is built until a solve), cached per mesh / dtype / device."""
    from .engine import Engine

    space = f.function_space()
    if space.local_rows is not None:
        raise ValueError("sample_function needs the whole field: gather a strip-sharded function first")
    mesh = space.mesh()
    key = (mesh.nx, mesh.ny, mesh.width, mesh.height, str(f.tensor.dtype), str(f.tensor.device))
    engine = _ENGINES.get(key)
    if engine is None:
        if len(_ENGINES) >= 4:
            _ENGINES.clear()
        engine = Engine(mesh.nx, mesh.ny, mesh.width, mesh.height, dtype=space.dtype_name,
                        device=f.tensor.device)
        _ENGINES[key] = engine
    return engine


def sample_grid(points: int, sample_type: str, N: int, domain_size):
    """Sample counts and spacing of ``sample_function`` (reference: FEM_src/utils.py:136-158):
    a multiple of N samples per unit length with at least ``points`` of them; "center" samples
    sit at (0.5 + i) / (multiplier N), "edges" at i / (multiplier N) with one more per axis.
    Returns (nsx, nsy, x0, dx, y0, dy, multiplier)."""
    if sample_type not in ("center", "edges"):
        raise ValueError(
            f"Unknown sample_type: {sample_type}. sample_type must be either 'center' or 'edges'")
    multiplier = int(np.ceil(points / N))
    nsx, nsy = (int(s * N * multiplier) for s in domain_size)
    step = 1.0 / (multiplier * N)
    if sample_type == "edges":
        return nsx + 1, nsy + 1, 0.0, step, 0.0, step, multiplier
    return nsx, nsy, 0.5 / (multiplier * N), step, 0.5 / (multiplier * N), step, multiplier


def sample_function(f: Function, points: int, sample_type: str, N: int | None = None,
                    domain_size=None):
    """Same arguments, sample positions and return value as the reference's ``sample_function``:
    ``(domain_rays, output_grid)`` with ``output_grid[yi, xi, :]`` = f at the sample point and
    ``output_grid.shape == (nsy, nsx, 1 or 2)``."""
    if not isinstance(f, Function):
        raise TypeError("sample_function expects a topomax_b200 Function")
    space = f.function_space()
    mesh = space.mesh()
    if N is None:
        N = mesh_to_N(mesh)
    if domain_size is None:
        domain_size = mesh_to_domain_size(mesh)
    nsx, nsy, x0, dx, y0, dy, _ = sample_grid(points, sample_type, N, domain_size)
    domain_rays = [np.linspace(0, s, ns) for s, ns in zip(domain_size, (nsx, nsy))]
    engine = _sampling_engine(f)
    grid = engine.sample_field(f.tensor, space.degree, nsx, nsy, x0, dx, y0, dy)
    return domain_rays, grid.cpu().numpy().astype(np.float64)
