"""Light structured-grid stand-ins for the dolfin objects the reference passes around
(``df.Mesh``, ``df.FunctionSpace``, ``df.Function``; reference: FEM_src/solver.py:38-55).

A ``Function`` wraps a torch CUDA tensor in the engine's layout; ``vector()[:]`` gives the
numpy get/set access the reference code uses (``rho.vector()[:] = values``).
"""
from __future__ import annotations

import numpy as np
import torch


class RectangleMesh:
    """``df.RectangleMesh(Point(0,0), Point(W,H), nx, ny)``, default "right" diagonal."""

    def __init__(self, width: float, height: float, nx: int, ny: int):
        self.width, self.height = float(width), float(height)
        self.nx, self.ny = int(nx), int(ny)

    @property
    def domain_size(self):
        return (self.width, self.height)

    def num_cells(self):
        return 2 * self.nx * self.ny

    def num_vertices(self):
        return (self.nx + 1) * (self.ny + 1)

    def hmin(self):
        # dolfin's cell diameter of a right triangle: its hypotenuse
        return float(np.hypot(self.width / self.nx, self.height / self.ny))


class FunctionSpace:
    """"CG" degree 1 (scalar, vertex grid) or vector "CG" degree 2 (half-step lattice)."""

    def __init__(self, mesh: RectangleMesh, family: str = "CG", degree: int = 1, *,
                 dtype: str = "float64", device=None, local_rows: tuple[int, int] | None = None):
        if family in ("TaylorHood", "TH"):  # vector P2 x P1 (the fluid problem's state), stored as [u | p]
            degree = (2, 1)
        elif family not in ("CG", "Lagrange", "P"):
            raise ValueError(f"unsupported element family {family!r}")
        if degree not in (1, 2, (2, 1)):
            raise ValueError("only P1 (scalar), vector-P2 and Taylor-Hood (vector P2 x P1) spaces exist on this path")
        self._mesh = mesh
        self.degree = degree
        self.dtype_name = dtype
        self.device = device
        # row-strip sharding: this rank stores cell rows [cl0, cl1) only (owned + halo rows)
        self.local_rows = local_rows
        ny_stored = mesh.ny if local_rows is None else local_rows[1] - local_rows[0]
        if degree == 1:
            self.shape = (ny_stored + 1, mesh.nx + 1)
        elif degree == 2:
            self.shape = (2 * ny_stored + 1, 2 * mesh.nx + 1, 2)
        else:  # flat [u | p]
            self.shape = (2 * (2 * ny_stored + 1) * (2 * mesh.nx + 1) + (ny_stored + 1) * (mesh.nx + 1),)

    def mesh(self):
        return self._mesh

    def dim(self):
        return int(np.prod(self.shape))


class _VectorView:
    def __init__(self, function: "Function"):
        self._f = function

    def __getitem__(self, key):
        return self._f.tensor.detach().cpu().numpy()[key]

    def __setitem__(self, key, values):
        t = self._f.tensor
        if isinstance(values, torch.Tensor):
            t[key] = values.to(device=t.device, dtype=t.dtype)
        elif np.isscalar(values):
            t[key] = float(values)
        else:
            t[key] = torch.as_tensor(np.asarray(values), dtype=t.dtype).to(t.device)

    def __len__(self):
        return self._f.tensor.numel()

    def get_local(self):
        return self[:]

    def set_local(self, values):
        self[:] = values


class Function:
    def __init__(self, space: FunctionSpace, tensor: torch.Tensor | None = None):
        self._space = space
        if tensor is None:
            dtype = torch.float64 if space.dtype_name == "float64" else torch.float32
            device = space.device
            if device is None:
                if not torch.cuda.is_available():
                    raise RuntimeError("topomax_b200 Functions live in CUDA memory; no GPU is available")
                device = torch.device("cuda", torch.cuda.current_device())
            tensor = torch.zeros(space.dim(), dtype=dtype, device=device)
        self.tensor = tensor

    def function_space(self):
        return self._space

    def vector(self):
        return _VectorView(self)

    def copy(self):
        return Function(self._space, self.tensor.clone())

    def grid(self) -> np.ndarray:
        """Values on the node grid: (ny+1, nx+1) for P1, (2ny+1, 2nx+1, 2) for P2."""
        return self.tensor.detach().cpu().numpy().reshape(self._space.shape)
