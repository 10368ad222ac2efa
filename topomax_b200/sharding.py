"""Host-side helpers of the row-strip sharded path (one process per GPU, torch.distributed for
the plumbing): slicing global fields into rank-local arrays (owned rows + halo rows) and
gathering the owned rows back.  The partition itself is computed by the library
(``tm_local_layout``); this module only mirrors its arithmetic for tests and I/O."""
from __future__ import annotations

import numpy as np
import torch


def partition_rows(ny: int, nranks: int, dist_levels: int):
    """Level-0 cell-row starts of every rank, as libtopomax_b200 computes them: strips aligned to
    2**dist_levels rows so that they nest across the sharded multigrid levels."""
    if nranks == 1:
        return [0, ny]
    align = 1 << dist_levels
    per = -(-(-(-ny // nranks)) // align) * align
    starts = [min(r * per, ny) for r in range(nranks)] + [ny]
    if starts[nranks - 1] >= ny:
        raise ValueError(f"mesh has too few cell rows ({ny}) for {nranks} ranks")
    return starts


def local_p1(engine, values_global: np.ndarray) -> torch.Tensor:
    """Rank-local P1 array (stored vertex rows cl0..cl1) of a global (ny+1, nx+1) field."""
    g = np.asarray(values_global).reshape(engine.ny + 1, engine.nx + 1)
    loc = np.ascontiguousarray(g[engine.cl0:engine.cl1 + 1])
    return torch.as_tensor(loc.ravel(), dtype=engine.dtype).to(engine.device)


def local_p2(engine, values_global: np.ndarray) -> torch.Tensor:
    g = np.asarray(values_global).reshape(2 * engine.ny + 1, 2 * engine.nx + 1, 2)
    loc = np.ascontiguousarray(g[2 * engine.cl0:2 * engine.cl1 + 1])
    return torch.as_tensor(loc.ravel(), dtype=engine.dtype).to(engine.device)


def _gather(engine, local: torch.Tensor, rows, row_shape, nrows_global, group=None) -> np.ndarray:
    lo, hi, goff = rows
    mine = local.detach().reshape(-1, *row_shape)[lo:hi].cpu().numpy()
    if engine.nranks == 1:
        return mine.reshape(-1)
    import torch.distributed as dist
    pieces = [None] * engine.nranks
    dist.all_gather_object(pieces, (goff, mine), group=group)
    out = np.empty((nrows_global, *row_shape), dtype=mine.dtype)
    for off, arr in pieces:
        out[off:off + arr.shape[0]] = arr
    return out.reshape(-1)


def gather_p1(engine, local: torch.Tensor, group=None) -> np.ndarray:
    """Global (ny+1)*(nx+1) array from the owned rows of every rank (returned on all ranks)."""
    return _gather(engine, local, engine.owned_p1_rows(), (engine.nx + 1,), engine.ny + 1, group)


def gather_p2(engine, local: torch.Tensor, group=None) -> np.ndarray:
    return _gather(engine, local, engine.owned_p2_rows(), (2 * engine.nx + 1, 2), 2 * engine.ny + 1, group)
