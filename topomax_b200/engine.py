"""Thin Python handle over the C ABI: torch is used for device memory and streams only."""
from __future__ import annotations

import ctypes
from ctypes import byref, c_double, c_int, c_void_p
from dataclasses import dataclass

import torch

from . import _lib
from .designs.definitions import Side

_SIDE_BITS = {
    Side.LEFT: _lib.SIDE_LEFT,
    Side.RIGHT: _lib.SIDE_RIGHT,
    Side.TOP: _lib.SIDE_TOP,
    Side.BOTTOM: _lib.SIDE_BOTTOM,
}

_TORCH_DTYPE = {"float64": torch.float64, "float32": torch.float32}
_TM_DTYPE = {"float64": _lib.TM_F64, "float32": _lib.TM_F32}


@dataclass
class SolveInfo:
    iterations: int
    relative_residual: float


def _require_cuda(device: torch.device):
    if not torch.cuda.is_available():
        raise RuntimeError(
            "topomax_b200 needs a CUDA device (sm_100a); torch.cuda.is_available() is False "
            "and there is no CPU fallback on this path."
        )
    if device.type != "cuda":
        raise RuntimeError(f"topomax_b200 runs on CUDA devices only, got '{device}'")


class Engine:
    """One structured mesh + material + filter radius bound to one GPU."""

    def __init__(self, nx: int, ny: int, width: float, height: float, *, lame_lambda: float = 1.0,
                 lame_mu: float = 1.0, simp_min: float = 1e-6, filter_radius: float = 0.0,
                 fixed_sides=(), dtype: str = "float64", device=None, rank: int = 0, nranks: int = 1,
                 dist_levels: int = 0):
        self.lib = _lib.load_library()
        self.device = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0) \
            if device is None else torch.device(device)
        _require_cuda(self.device)
        if dtype not in _TM_DTYPE:
            raise ValueError(f"dtype must be 'float64' or 'float32', got {dtype!r}")
        self.dtype_name = dtype
        self.dtype = _TORCH_DTYPE[dtype]
        self.nx, self.ny = int(nx), int(ny)
        self.width, self.height = float(width), float(height)
        self.rank, self.nranks = int(rank), int(nranks)
        bits = 0
        for side in fixed_sides:
            if side not in _SIDE_BITS:
                raise ValueError(f"Malformed side: {side}")
            bits |= _SIDE_BITS[side]
        cfg = _lib.TmConfig(
            nx=self.nx, ny=self.ny, width=self.width, height=self.height,
            lame_lambda=lame_lambda, lame_mu=lame_mu, simp_min=simp_min,
            filter_radius=filter_radius, fixed_sides=bits, dtype=_TM_DTYPE[dtype],
            device=self.device.index or 0, rank=self.rank, nranks=self.nranks, mg_dist_levels=int(dist_levels),
        )
        handle = c_void_p()
        _lib.check(self.lib.tm_create(byref(cfg), byref(handle)))
        self._h = handle
        # rank-local storage: cell rows [cl0, cl1) are stored (owned [c0, c1) + halo rows)
        lay = (c_int * 10)()
        _lib.check(self.lib.tm_local_layout(self._h, lay, 10))
        (_, _, _, _, self.cl0, self.cl1, self.c0, self.c1, owns_top, self.dist_levels) = tuple(lay)
        self.owns_top = bool(owns_top)
        self.ny_local = self.cl1 - self.cl0
        self.n1 = (self.nx + 1) * (self.ny_local + 1)
        self.n2 = (2 * self.nx + 1) * (2 * self.ny_local + 1)
        self.nu = 2 * self.n2
        self._sync_stream()

    # ---------------------------------------------------------------- sharding
    def init_comm(self, group=None):
        """Create the engine's NCCL communicator; the unique id travels over torch.distributed."""
        if self.nranks == 1:
            return
        import torch.distributed as dist
        buf = ctypes.create_string_buffer(128)
        if self.rank == 0:
            _lib.check(self.lib.tm_comm_unique_id(buf))
        on_gpu = dist.get_backend(group) == "nccl"
        t = torch.tensor(list(buf.raw), dtype=torch.uint8, device=self.device if on_gpu else "cpu")
        dist.broadcast(t, src=0, group=group)
        raw = bytes(t.cpu().tolist())
        self._sync_stream()
        _lib.check(self.lib.tm_comm_init(self._h, raw))

    @property
    def peer_memory_active(self) -> bool:
        """True when halo rows / scalar sums go through the library's peer-memory kernels
        (``TM_OPT_P2P`` or ``TM_P2P=1``) rather than NCCL calls."""
        lay = (c_int * 11)()
        _lib.check(self.lib.tm_local_layout(self._h, lay, 11))
        return bool(lay[10])

    def owned_p1_rows(self):
        """(first, last+1) local vertex rows this rank owns, and their global offset."""
        lo = self.c0 - self.cl0
        hi = self.c1 - self.cl0 + (1 if self.owns_top else 0)
        return lo, hi, self.c0

    def owned_p2_rows(self):
        lo = 2 * (self.c0 - self.cl0)
        hi = 2 * (self.c1 - self.cl0) + (1 if self.owns_top else 0)
        return lo, hi, 2 * self.c0

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self.lib.tm_destroy(h)
            except Exception:
                pass
            self._h = None

    # ---------------------------------------------------------------- plumbing
    def _sync_stream(self):
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.tm_set_stream(self._h, c_void_p(stream)))

    def _p(self, t: torch.Tensor, n: int):
        if t.device != self.device or t.dtype != self.dtype or not t.is_contiguous() or t.numel() != n:
            raise ValueError(
                f"expected a contiguous {self.dtype} tensor of {n} elements on {self.device}, "
                f"got {t.dtype} {tuple(t.shape)} on {t.device}"
            )
        return c_void_p(t.data_ptr())

    def empty_p1(self):
        return torch.empty(self.n1, dtype=self.dtype, device=self.device)

    def empty_p2(self):
        return torch.empty(self.nu, dtype=self.dtype, device=self.device)

    def set_option(self, option: int, value: float):
        _lib.check(self.lib.tm_set_option(self._h, option, float(value)))

    # ---------------------------------------------------------------- operations
    def load_vector(self, body_force, tractions) -> torch.Tensor:
        """body_force: designs.definitions.Force | None; tractions: list[Traction] | None."""
        spec = _lib.TmLoads()
        if body_force is not None:
            spec.has_force = 1
            spec.force_center[0], spec.force_center[1] = body_force.region.center
            spec.force_radius = body_force.region.radius
            spec.force_value[0], spec.force_value[1] = body_force.value
        tractions = tractions or []
        if len(tractions) > 8:
            raise ValueError("at most 8 tractions are supported")
        spec.ntractions = len(tractions)
        for k, t in enumerate(tractions):
            if t.side not in _SIDE_BITS:
                raise ValueError(f"Malformed side: {t.side}")
            spec.traction_side[k] = _SIDE_BITS[t.side]
            spec.traction_center[k] = t.center
            spec.traction_length[k] = t.length
            spec.traction_value[k][0], spec.traction_value[k][1] = t.value
        b = self.empty_p2()
        self._sync_stream()
        _lib.check(self.lib.tm_load_vector(self._h, byref(spec), self._p(b, self.nu)))
        return b

    def filter_apply(self, values: torch.Tensor, *, assembled: bool, rtol=1e-12, maxit=20000,
                     out: torch.Tensor | None = None):
        out = self.empty_p1() if out is None else out
        iters, relres = c_int(0), c_double(0.0)
        self._sync_stream()
        _lib.check(self.lib.tm_filter_apply(
            self._h, 1 if assembled else 0, self._p(values, self.n1), self._p(out, self.n1),
            rtol, maxit, byref(iters), byref(relres)))
        return out, SolveInfo(iters.value, relres.value)

    def elast_matvec(self, xi, x, penalty=3.0, out=None):
        out = self.empty_p2() if out is None else out
        self._sync_stream()
        _lib.check(self.lib.tm_elast_matvec(self._h, self._p(xi, self.n1), penalty,
                                            self._p(x, self.nu), self._p(out, self.nu)))
        return out

    def elast_diag_inverse(self, xi, penalty=3.0):
        out = self.empty_p2()
        self._sync_stream()
        _lib.check(self.lib.tm_elast_diag(self._h, self._p(xi, self.n1), penalty, self._p(out, self.nu)))
        return out

    def state_solve(self, xi, b, penalty=3.0, *, rtol=1e-10, maxit=200000, u=None, warm_start=False):
        if u is None:
            u = self.empty_p2()
            warm_start = False
        iters, relres = c_int(0), c_double(0.0)
        self._sync_stream()
        _lib.check(self.lib.tm_state_solve(
            self._h, self._p(xi, self.n1), penalty, self._p(b, self.nu), self._p(u, self.nu),
            rtol, maxit, 1 if warm_start else 0, byref(iters), byref(relres)))
        return u, SolveInfo(iters.value, relres.value)

    def dot_p2(self, u, b) -> float:
        out = c_double(0.0)
        self._sync_stream()
        _lib.check(self.lib.tm_dot_p2(self._h, self._p(u, self.nu), self._p(b, self.nu), byref(out)))
        return out.value

    def sens_rhs(self, xi, u, penalty=3.0):
        out = self.empty_p1()
        self._sync_stream()
        _lib.check(self.lib.tm_sens_rhs(self._h, self._p(xi, self.n1), penalty, self._p(u, self.nu),
                                        self._p(out, self.n1)))
        return out

    def md_halfstep(self, psi, grad, alpha, out=None):
        out = self.empty_p1() if out is None else out
        self._sync_stream()
        _lib.check(self.lib.tm_md_halfstep(self._h, self._p(psi, self.n1), self._p(grad, self.n1),
                                           float(alpha), self._p(out, self.n1)))
        return out

    def md_volume(self, half, c: float):
        vol, dvol = c_double(0.0), c_double(0.0)
        self._sync_stream()
        _lib.check(self.lib.tm_md_volume(self._h, self._p(half, self.n1), float(c), byref(vol), byref(dvol)))
        return vol.value, dvol.value

    def md_project(self, half, volume: float, tol: float = 1e-12, maxit: int = 50):
        """Device-resident Newton iteration for the volume shift (``tm_md_project``): returns
        ``(c, iterations, status)``, status 1 = converged, 2 = zero derivative, 0 = not converged."""
        c, iters, status = c_double(0.0), c_int(0), c_int(0)
        self._sync_stream()
        _lib.check(self.lib.tm_md_project(self._h, self._p(half, self.n1), float(volume), float(tol), int(maxit),
                                          byref(c), byref(iters), byref(status)))
        return c.value, iters.value, status.value

    def md_apply(self, half, c: float, psi_prev, psi_out, rho_out):
        dsq, vol = c_double(0.0), c_double(0.0)
        self._sync_stream()
        _lib.check(self.lib.tm_md_apply(self._h, self._p(half, self.n1), float(c),
                                        self._p(psi_prev, self.n1), self._p(psi_out, self.n1),
                                        self._p(rho_out, self.n1), byref(dsq), byref(vol)))
        return dsq.value, vol.value

    def integrate(self, values) -> float:
        out = c_double(0.0)
        self._sync_stream()
        _lib.check(self.lib.tm_integrate(self._h, self._p(values, self.n1), byref(out)))
        return out.value

    def sample_field(self, field, degree: int, nsx: int, nsy: int, x0: float, dx: float, y0: float,
                     dy: float):
        """(nsy, nsx, degree) tensor of field values at (x0 + sx dx, y0 + sy dy); ``field`` is a
        P1 (degree 1) or vector-P2 (degree 2) tensor of this engine's mesh."""
        if degree not in (1, 2):
            raise ValueError("degree must be 1 (P1) or 2 (vector P2)")
        out = torch.empty((int(nsy), int(nsx), degree), dtype=self.dtype, device=self.device)
        _lib.check(self.lib.tm_sample_field(self._h, degree, self._p(field, self.n1 if degree == 1 else self.nu),
                                            int(nsx), int(nsy), float(x0), float(dx), float(y0), float(dy),
                                            out.data_ptr()))
        return out

    def last_solve_stats(self) -> dict:
        buf = (c_double * 17)()
        _lib.check(self.lib.tm_last_solve_stats(self._h, buf, 17))
        return {"iterations": int(buf[0]), "vcycles": int(buf[1]), "fine_applies": int(buf[2]),
                "levels": int(buf[3]), "lambda_max": buf[4],
                "fine_launches_total": {"plain": int(buf[5]), "dot": int(buf[6]), "resid": int(buf[7]),
                                        "cheb": int(buf[8])},
                "tail_first_level": int(buf[9]), "tail_cluster": int(buf[10]),
                "warm_start_used": bool(buf[11]),
                # multigrid levels cycled `cycles` times per visit of their parent (W-cycle window; V-cycle: [-1, -1, 1])
                "cycle_window": [int(buf[12]), int(buf[13]), int(buf[14])],
                # estimate of the relative residual fp arithmetic cannot resolve (0.0: not estimated: no initial
                # guess, or option 137 = 0) and the tolerance the PCG stopped at, max(rtol, 0.5 x that estimate)
                "fp_floor_estimate": buf[15], "rtol_used": buf[16]}

    def profile_read(self) -> dict:
        """Milliseconds / launch counts of the fine-level operator kernel per epilogue since the
        last read (needs ``set_option(OPT_PROFILE, 1)``)."""
        buf = (c_double * 136)()
        _lib.check(self.lib.tm_profile_read(self._h, buf, 136))
        names = ("plain", "dot", "resid", "cheb")
        out = {n: {"ms": buf[i], "launches": int(buf[4 + i])} for i, n in enumerate(names)}
        levels = []
        for lvl in range(16):
            base = 8 + 8 * lvl
            ms, cnt = sum(buf[base:base + 4]), int(sum(buf[base + 4:base + 8]))
            if cnt:
                levels.append({"level": lvl, "ms": ms, "launches": cnt})
        self.last_level_profile = levels
        return out

    def launch_count(self) -> int:
        return int(self.lib.tm_launch_count())

    LEDGER_CATEGORIES = (
        ["fine_op_plain", "fine_op_dot", "fine_op_resid", "fine_op_cheb", "fine_op_resid0", "fine_op_chebdot"]
        + [f"level{lvl}_op" for lvl in range(1, 15)]
        + ["restrict", "prolong", "cheb_first", "pcg_update", "pcg_direction", "reductions", "copies", "tail",
           "coarse_solve", "filter", "mirror_descent", "sensitivity", "setup_moments", "setup_diag", "setup_eig",
           "halo", "allreduce", "gather", "convert", "other"])

    def ledger_read(self, reset: bool = True) -> dict:
        """Per category: algorithmic bytes, launches and (``OPT_PROFILE`` = 3 only) milliseconds since the
        last reset -- ``tm_ledger_read``."""
        c = _lib.LEDGER_CATEGORIES
        buf = (c_double * (3 * c))()
        _lib.check(self.lib.tm_ledger_read(self._h, buf, 3 * c, 1 if reset else 0))
        out = {}
        for i, name in enumerate(self.LEDGER_CATEGORIES):
            if buf[c + i]:
                out[name] = {"bytes": buf[i], "launches": int(buf[c + i]), "ms": buf[2 * c + i]}
        return out

    # ---------------------------------------------------------------- multigrid diagnostics
    def mg_levels(self):
        """[(nx, ny, dl, dr, db, dt)] per level of the hierarchy."""
        info, nl = (c_int * 6)(), c_int(0)
        _lib.check(self.lib.tm_mg_level_info(self._h, 0, info, byref(nl)))
        out = []
        for lvl in range(nl.value):
            _lib.check(self.lib.tm_mg_level_info(self._h, lvl, info, byref(nl)))
            out.append(tuple(info))
        return out

    def mg_debug(self, xi, op: int, level: int, vec: torch.Tensor, out_size: int):
        out = torch.zeros(out_size, dtype=self.dtype, device=self.device)
        self._sync_stream()
        _lib.check(self.lib.tm_mg_debug(self._h, self._p(xi, self.n1), op, level,
                                        c_void_p(vec.data_ptr()), c_void_p(out.data_ptr())))
        return out
