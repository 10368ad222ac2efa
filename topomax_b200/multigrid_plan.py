"""Host-side mirror of the multigrid planning arithmetic of libtopomax_b200 (``Engine::plan_levels``,
``plan_tail``, ``repeats``, ``level_degree`` in csrc/tm_engine.cu) -- for documentation, tests and tools only:
the library plans for itself and reports what it chose through ``tm_mg_level_info`` / ``tm_last_solve_stats``.

The reference has no counterpart (it factorises with MUMPS: FEM_src/pde_solver.py:130-131); this is the
shape of the preconditioner that replaces the factorisation."""
from __future__ import annotations

from dataclasses import dataclass

TAIL_MAX_NODES = 2304      # levels of at most this many lattice nodes run inside the cluster tail kernel
TAIL_MAX_LEVELS = 8
BIG_MESH_DOFS = 1 << 24    # from here on a visit of the small levels is noise next to the fine level


@dataclass(frozen=True)
class Level:
    index: int
    nx: int
    ny: int
    cycles: int           # visits per visit of the parent level (1 = V-cycle, 2 = the W window)
    smoothing_steps: int  # Chebyshev-Jacobi steps before and after the coarse-grid correction (0: exact solve)
    in_tail: bool         # handled inside tail_vcycle_kernel

    @property
    def lattice_nodes(self) -> int:
        return (2 * self.nx + 1) * (2 * self.ny + 1)


def level_sizes(nx: int, ny: int, coarse_cells: int = 4):
    """Cell counts per level: ceil-halving until max(nx, ny) <= coarse_cells (or 1 x 1)."""
    sizes = [(nx, ny)]
    while max(nx, ny) > coarse_cells:
        nxc, nyc = (nx + 1) // 2, (ny + 1) // 2
        if (nxc, nyc) == (nx, ny):
            break
        nx, ny = nxc, nyc
        sizes.append((nx, ny))
    return sizes


def plan(nx: int, ny: int, *, coarse_cells: int = 4, fine_steps: int = 1, coarse_steps: int = 3,
         first_tail_candidate: int = 1):
    """Levels of the elasticity multigrid as the library plans them by default (single GPU, or the replicated
    levels of a sharded engine when ``first_tail_candidate`` = its number of sharded levels)."""
    sizes = level_sizes(nx, ny, coarse_cells)
    nl = len(sizes)
    big = 2 * (2 * nx + 1) * (2 * ny + 1) >= BIG_MESH_DOFS
    lo, hi = (4, 32) if big else (8, 16)
    cycles = [1] * nl
    for l in range(1, nl - 1):  # never the finest level, never the exactly solved coarsest one
        if lo <= min(sizes[l]) <= hi:
            cycles[l] = 2
    tail_first = -1
    for l in range(max(1, first_tail_candidate), nl):
        if (2 * sizes[l][0] + 1) * (2 * sizes[l][1] + 1) <= TAIL_MAX_NODES:
            tail_first = l
            break
    if tail_first < 0 or nl - tail_first < 2:
        tail_first = -1
    else:
        tail_first = max(tail_first, nl - TAIL_MAX_LEVELS)
    light = 2 if any(c > 1 for c in cycles) else 0
    if tail_first >= 0:
        light = min(light, tail_first - 1)
    out = []
    for l, (lx, ly) in enumerate(sizes):
        if l == nl - 1 and nl > 1:
            steps = 0
        elif l == 0:
            steps = fine_steps
        else:
            steps = min(2, coarse_steps) if l <= light else coarse_steps
        out.append(Level(l, lx, ly, cycles[l], steps, tail_first >= 0 and l >= tail_first))
    return out


def cycle_window(levels):
    """[first, last, cycles] as ``tm_last_solve_stats`` reports it ([-1, -1, 1]: plain V-cycle)."""
    idx = [lv.index for lv in levels if lv.cycles > 1]
    return [idx[0], idx[-1], max(lv.cycles for lv in levels)] if idx else [-1, -1, 1]


def visits(levels):
    """How often each level is visited per application of the preconditioner."""
    v, out = 1, []
    for lv in levels:
        v *= lv.cycles
        out.append(v)
    return out
