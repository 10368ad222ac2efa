/* libtopomax_b200 -- C ABI of the B200-native FEM elasticity inner loop of topomax.
 *
 * The reference (Emilinya/topomax) has no FFI: its seam is the Python classes
 * FEMSolver / ElasticityProblem / HelmholtzFilter, which call dolfin.  This header is the
 * boundary a binding for that seam links against; every entry point names the reference
 * code whose arithmetic it replaces.  Conventions:
 *   - plain C types only; every array argument is a CUDA DEVICE pointer owned by the caller
 *     (the Python host passes torch.Tensor.data_ptr()); element type = the engine dtype;
 *   - the engine owns only scratch memory; no ownership is transferred; on a sharded engine
 *     the non-const field arguments (xi, x, in) get their halo rows refreshed in place;
 *   - every function returns 0 on success, a negative TM_ERR_* code otherwise, and
 *     tm_last_error() returns the message of the last failure on the calling thread;
 *   - work is enqueued on the stream given to tm_set_stream (default: the legacy stream);
 *     functions that return scalars synchronise that stream before returning.
 *
 * Layouts:
 *   P1 fields (rho, psi, xi, G):  (ny+1) x (nx+1) vertex grid, row-major, v = iy*(nx+1)+ix
 *   P2 fields (u, b):  (2ny+1) x (2nx+1) half-step lattice, row-major, 2 interleaved
 *                      components per node, dof = 2*(j*(2nx+1)+i) + comp
 */
#ifndef TOPOMAX_B200_H
#define TOPOMAX_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tm_engine_s* tm_handle;

enum { TM_OK = 0, TM_ERR_INVALID = -1, TM_ERR_UNSUPPORTED = -2, TM_ERR_CUDA = -3,
       TM_ERR_NOT_CONVERGED = -4 };

enum { TM_F64 = 0, TM_F32 = 1 };
enum { TM_SIDE_LEFT = 1, TM_SIDE_RIGHT = 2, TM_SIDE_TOP = 4, TM_SIDE_BOTTOM = 8 };
enum { TM_PRECOND_JACOBI = 0, TM_PRECOND_MULTIGRID = 1 };

/* integer / real options for tm_set_option */
enum {
    TM_OPT_PRECOND = 1,      /* TM_PRECOND_*                      (default MULTIGRID)   */
    TM_OPT_CHEB_DEGREE = 2,  /* Chebyshev-Jacobi smoothing steps on the finest level (default 1; levels 1-2: 2 when a cycle
                                window is active, the other coarse levels 3) */
    TM_OPT_CHECK_EVERY = 3,  /* PCG iterations between residual read-backs (default 0 = automatic: every iteration with
                                the multigrid preconditioner, from two iterations before the count of this engine's
                                previous warm-started solve; every 25 with Jacobi) */
    TM_OPT_MG_COARSE_CELLS = 4, /* stop coarsening at max(nx,ny) <= this (1..4, default 4) */
    TM_OPT_PROFILE = 5,      /* 1: time every fine-level operator launch with CUDA events; 2: the operator
                                launches of every multigrid level; 3: every launch, by ledger category
                                (tm_ledger_read).  2 and 3 switch CUDA-graph replay off: diagnostics only */
    TM_OPT_CYCLE_FIRST = 133, /* multigrid cycle: levels [CYCLE_FIRST, CYCLE_LAST] are cycled CYCLE_GAMMA times per  */
    TM_OPT_CYCLE_LAST = 134,  /* visit of their parent (a W-cycle on that window, a V-cycle elsewhere; pre- and post-  */
    TM_OPT_CYCLE_GAMMA = 135, /* smoother are the same polynomial, so the cycle stays symmetric).  GAMMA 0 (default):  */
                              /* automatic window = the levels whose short side has 8..16 cells (4..32 on meshes of    */
                              /* >= 2^24 dofs); 1: plain V-cycle; 2..4: explicit window                               */
    TM_OPT_FP_FLOOR_FACTOR = 137, /* state solves that start from an initial guess stop at max(rtol, factor x floor),     */
                              /* floor = ||K (u o delta)|| / ||b||, delta_i = +-eps/2: the relative residual fp arithmetic */
                              /* cannot resolve on this mesh, estimated by one operator pass (default 0.5; 0 = off:       */
                              /* iterate until the RECURSIVE residual meets rtol).  tm_last_solve_stats [15], [16]        */
    TM_OPT_P2P = 130         /* sharded runs: halo exchange and scalar all-reduce by our own kernels over
                                peer-mapped memory (NVLink) instead of NCCL calls; collective, set alike
                                on every rank (default 1 since round 2; env TM_P2P=0 selects NCCL,
                                which is also the fall-back when a peer window cannot be mapped)   */
};

/* Mesh, material and filter of one problem.
 * reference: FEM_src/solver.py:38-49 (RectangleMesh((0,0),(W,H),nx,ny), "right" diagonal),
 * FEM_src/elasisity_problem.py:93-99 (Lame constants), src/penalizers.py:32-34 (minimum),
 * FEM_src/filter.py:15-38 (epsilon), FEM_src/elasisity_problem.py:183-192 (fixed sides). */
typedef struct {
    int nx, ny;
    double width, height;
    double lame_lambda, lame_mu;
    double simp_min;
    double filter_radius;
    int fixed_sides; /* TM_SIDE_* bitmask */
    int dtype;       /* TM_F64 / TM_F32 */
    int device;      /* CUDA device ordinal */
    /* row-strip sharding, one process per GPU: this engine is rank `rank` of
     * `nranks` (1 = unsharded).  nx, ny, width, height always describe the GLOBAL mesh; the
     * library partitions the cell rows itself (tm_local_layout) and every array argument of a
     * sharded engine is the rank-LOCAL array (owned rows + halo rows).  mg_dist_levels: how
     * many multigrid levels stay sharded (0 = automatic). */
    int rank, nranks, mg_dist_levels;
} tm_config;

/* reference: designs/definitions.py Force / Traction, evaluated by BodyForce.eval and
 * TractionExpression.eval (FEM_src/elasisity_problem.py:26-33, :47-70). */
typedef struct {
    int has_force;
    double force_center[2], force_radius, force_value[2];
    int ntractions;         /* <= 8 */
    int traction_side[8];   /* one TM_SIDE_* each */
    double traction_center[8], traction_length[8], traction_value[8][2];
} tm_loads;

int tm_create(const tm_config* cfg, tm_handle* out);
int tm_destroy(tm_handle h);
int tm_set_stream(tm_handle h, void* cuda_stream);
int tm_set_option(tm_handle h, int option, double value);
const char* tm_last_error(void);
const char* tm_version(void);

/* Sharding plumbing.  Rank 0 calls tm_comm_unique_id (an ncclUniqueId, 128 bytes), the host
 * broadcasts it (torch.distributed), every rank calls tm_comm_init: the engine then owns an
 * NCCL communicator for halo rows (ncclSend/Recv), dot products (ncclAllReduce) and the gather
 * of the first replicated multigrid level (ncclBroadcast).
 * By default (TM_OPT_P2P = 1; env TM_P2P=0 or the option turn it off) the halo rows and the scalar sums
 * travel instead through peer-mapped windows (cudaIpc) written and polled by the library's own kernels
 * over NVLink (csrc/tm_p2p.cuh); NCCL then only carries the window handles and the coarse-level gather.
 * tm_local_layout: out = {rank, nranks, nx, ny_global, cl0, cl1, c0, c1, owns_top, dist_levels,
 * peer_memory_active}: this rank stores cell rows [cl0, cl1) and owns [c0, c1); local P1 arrays are
 * (cl1-cl0+1) x (nx+1), local P2 arrays (2(cl1-cl0)+1) x (2nx+1) x 2. */
int tm_comm_unique_id(char* id128);
int tm_comm_init(tm_handle h, const char* id128);
int tm_local_layout(tm_handle h, int* out, int n);

/* b = int f_h.v dx + int t_h.v ds (not yet zeroed on the fixed sides).
 * reference: l_func, FEM_src/elasisity_problem.py:120-124, assembled once by
 * SmartMumpsSolver.__init__ (FEM_src/pde_solver.py:102-104). */
int tm_load_vector(tm_handle h, const tm_loads* loads, void* b);

/* Helmholtz filter  (eps^2 K1 + M1) out = rhs.
 * rhs_kind 0: `in` is a nodal P1 function, rhs = M1 in   (HelmholtzFilter.apply on a Function)
 * rhs_kind 1: `in` is an assembled right-hand side       (apply on a UFL expression)
 * reference: FEM_src/filter.py:27-41 + SmartMumpsSolver.solve, FEM_src/pde_solver.py:106-133. */
int tm_filter_apply(tm_handle h, int rhs_kind, void* in, void* out, double rtol, int maxit,
                    int* iters, double* relres);

/* y = K(xi) x with identity rows on the fixed sides, r(xi) = m + (1-m) xi^penalty.
 * penalty = 3 (every elasticity design of the reference) is the fast path: closed-form SIMP
 * moments inside the fine-level operator.  Any other penalty > 0 is supported through level-0
 * moments stored once per call (integer penalty <= 16: exact; otherwise a 16-point degree-6 rule,
 * parity unpinned); penalty <= 0 -> TM_ERR_INVALID.  The same holds for tm_elast_diag,
 * tm_state_solve and tm_sens_rhs (reference: src/penalizers.py:36-46, penalties loop
 * src/solver.py:230-231).
 * reference: a_func, FEM_src/elasisity_problem.py:112-118 assembled in
 * FEM_src/pde_solver.py:117-119, bc.apply :125. */
int tm_elast_matvec(tm_handle h, void* xi, double penalty, void* x, void* y);
/* dinv = 1 / diag K(xi) (1 on fixed dofs) */
int tm_elast_diag(tm_handle h, void* xi, double penalty, void* dinv);

/* u = K(xi)^-1 b with u = 0 on the fixed sides, by preconditioned CG to ||r|| <= rtol ||b||.
 * flags bit 0: use the incoming u as initial guess.  With an initial guess the iteration also stops once the
 * recursive residual is below half the level fp arithmetic can resolve on this mesh (TM_OPT_FP_FLOOR_FACTOR):
 * at 2e8 dofs that level is 2.6e-9 in fp64, i.e. a direct solver's residual (what the reference's MUMPS returns);
 * *relres then reports the recursive residual at the stop, and tm_last_solve_stats the floor and the tolerance used.
 * reference: ElasticityProblem.forward, FEM_src/elasisity_problem.py:168-169 ->
 * SmartMumpsSolver.solve, FEM_src/pde_solver.py:106-133 (LUSolver("mumps")). */
int tm_state_solve(tm_handle h, void* xi, double penalty, const void* b, void* u,
                   double rtol, int maxit, int flags, int* iters, double* relres);

/* *out = u . b  (compliance; reference: FEM_src/elasisity_problem.py:161-164) */
int tm_dot_p2(tm_handle h, const void* u, const void* b, double* out);

/* out_i = int -r'(xi_h)(lambda (div u)^2 + 2 mu eps(u):eps(u)) phi_i dx  (assembled P1 rhs)
 * reference: calculate_objective_gradient, FEM_src/elasisity_problem.py:146-150 */
int tm_sens_rhs(tm_handle h, const void* xi, double penalty, const void* u, void* out);

/* half = psi - alpha G            reference: Solver.step, src/solver.py:191-192 */
int tm_md_halfstep(tm_handle h, const void* psi, const void* grad, double alpha, void* half);
/* vol = int expit(half+c), dvol = int expit'(half+c)   reference: Solver.project,
 * src/solver.py:158-162 with FEMSolver.integrate, FEM_src/solver.py:81-84 */
int tm_md_volume(tm_handle h, const void* half, double c, double* vol, double* dvol);
/* Newton iteration for the volume shift c with int expit(half + c) = volume, the iterate resident on the
 * device: scipy.optimize.newton(error, 0, fprime, tol, maxiter) of Solver.project, src/solver.py:166-174
 * (c0 = 0, step p = c - f/f', converged when |p - c| <= tol), batches of iterates enqueued without a host
 * round trip.  *status: 1 converged, 2 zero derivative, 0 not converged within maxit -- in the last two
 * cases the caller falls back to the bracketing search of src/solver.py:175-186 through tm_md_volume. */
int tm_md_project(tm_handle h, const void* half, double volume, double tol, int maxit, double* c, int* iters,
                  int* status);
/* psi = half + c, rho = expit(psi), *delta_sq = int (rho - expit(psi_prev))^2, *vol = int rho
 * reference: src/solver.py:186,262,286-288 */
int tm_md_apply(tm_handle h, const void* half, double c, const void* psi_prev, void* psi,
                void* rho, double* delta_sq, double* vol);
/* *out = int values dx (nodal quadrature)   reference: FEM_src/solver.py:81-84 */
int tm_integrate(tm_handle h, const void* values, double* out);

/* out[sy][sx][c] = field(x0 + sx dx, y0 + sy dy): point evaluation of a P1 (degree 1, one
 * component) or vector-P2 (degree 2, two components) field on a regular grid of nsx x nsy sample
 * points inside the rectangle; `out` is a device array of nsx*nsy*degree values of the engine's
 * dtype.  Unsharded engines only.
 * reference: the  f(x, y)  evaluations of sample_function, FEM_src/utils.py:112-162 (consumed
 * by plot.py:54-69) */
int tm_sample_field(tm_handle h, int degree, const void* field, int nsx, int nsy, double x0, double dx,
                    double y0, double dy, void* out);

/* statistics of the last tm_state_solve: out[0] iterations, [1] V-cycles, [2] fine-level
 * operator applications, [3] levels, [4] lambda_max estimate of level 0, [5..8] cumulative
 * level-0 operator launches per epilogue, [9] first level of the cluster tail (-1: none), [10] its
 * cluster size, [11] 1 if the caller's initial guess was kept (a warm start whose residual
 * exceeds that of the zero guess is dropped), [12..14] first / last multigrid level cycled more than
 * once per visit of its parent and the number of cycles (-1, -1, 1: plain V-cycle), [15] the relative
 * fp floor estimated for this solve (0: none), [16] the tolerance the iteration stopped at */
int tm_last_solve_stats(tm_handle h, double* out, int n);

/* Measurement support (bench.py): with TM_OPT_PROFILE on, out[0..3] = milliseconds spent in the
 * fine-level operator kernel per epilogue (plain, dot, residual, Chebyshev) since the last
 * read, out[4..7] = launch counts.  tm_launch_count: kernels launched by the library so far. */
int tm_profile_read(tm_handle h, double* out, int n);
long long tm_launch_count(void);

/* Measurement ledger (bench.py's step-level roofline, SURVEY.md section 8d "per-step figure = sum of
 * bytes / sum of time"): every launch and collective of the engine files its ALGORITHMIC bytes
 * (unique operand bytes read + written) under one of TM_LEDGER_CATEGORIES categories; launches
 * replayed from a CUDA graph are filed with the totals recorded at capture.
 *   out[0 .. C)        bytes per category since the last reset
 *   out[C .. 2C)       launches (kernels, copies, collectives) per category
 *   out[2C .. 3C)      milliseconds per category -- only with TM_OPT_PROFILE = 3, where every launch
 *                      is followed by an event record and the gap between two records goes to the
 *                      later one (launch gaps included: what the step really spends there)
 * Categories: 0-5 level-0 operator by epilogue (plain, p.Ap, residual, Chebyshev, fused first step +
 * residual, Chebyshev + r.z); 6-19 operator of multigrid level 1..14; 20 restriction; 21 prolongation;
 * 22 first smoothing step; 23 PCG update; 24 PCG direction; 25 dots / norms; 26 copies, masks;
 * 27 cluster tail; 28 coarsest solve; 29 filter; 30 mirror descent; 31 sensitivity; 32-34 hierarchy
 * set-up (moments, diagonals, smoother bounds); 35 halo exchange; 36 all-reduce; 37 gather;
 * 38 precision conversion; 39 other.  reset != 0 clears the ledger after reading. */
enum { TM_LEDGER_CATEGORIES = 40 };
int tm_ledger_read(tm_handle h, double* out, int n, int reset);

/* Diagnostics for the parity tests: the multigrid hierarchy built for xi.
 *   op 0: out = A_level in          op 1: out(level) = P in(level+1)
 *   op 2: out(level+1) = P^T in(level)   op 3: out = V-cycle(in) on level 0
 *   op 4: out = A_coarsest^-1 in    op 5: out = inverse diagonal of the level
 * tm_mg_level_info: info6 = {nx, ny, dl, dr, db, dt} of `level`; *nlevels = level count. */
int tm_mg_debug(tm_handle h, void* xi, int op, int level, const void* in, void* out);
int tm_mg_level_info(tm_handle h, int level, int* info6, int* nlevels);

/* Loop-back check of the peer-memory kernels on ONE GPU: `nranks` streams of this process play the
 * ranks (windows = plain allocations of the current device), run `epochs` halo exchanges of
 * 4 rows up / 3 rows down of `row_elems` doubles without host synchronisation, plus a 2-value
 * all-reduce every `reduce_every` epochs.  report4 = {data mismatches, ranks with a timed-out poll,
 * final halo epoch, final reduction epoch}. */
int tm_p2p_selftest(int nranks, int epochs, int row_elems, int reduce_every, double* report4);

/* the reference's FEM fluid problem (FEM_src/fluid_problem.py) on the same mesh:
 * Taylor-Hood (vector-P2 velocity on the half-step lattice, P1 pressure on the vertices)
 * Stokes-Brinkman state equation with velocities prescribed on the whole boundary.  fp64, one GPU.
 * Vectors: velocity [Ly][Lx][2] like the elasticity displacement; "up" = [velocity | pressure],
 * 2*Lx*Ly + (nx+1)(ny+1) doubles.  All pointers are device pointers.
 *   tm_fluid_create       FluidProblem.__init__ / create_solution_space (:50-66,150-153);
 *                         r_min, r_max: FluidPenalizer (src/penalizers.py:52-55)
 *   tm_fluid_set_density  set_penalization(q) + the rho-dependent part of a_func (:76-86): weighted
 *                         mass matrices (FIAT 12-point rule) and the MINRES preconditioner
 *   tm_fluid_state_solve  forward(rho) = SmartMumpsSolver.solve(a_arg=rho) (:124-125):
 *                         boundary_velocity holds the Dirichlet values on boundary nodes (zero
 *                         inside); preconditioned MINRES on the symmetric saddle-point system to
 *                         rtol on the preconditioned residual; the pressure comes out with an
 *                         arbitrary constant (the reference's matrix is singular)
 *   tm_fluid_objective    calculate_objective's assemble(0.5*(r u^2 + mu grad(u)^2)*dx) (:114-122)
 *   tm_fluid_sens_rhs     right-hand side of df.project(0.5 r'(rho) u^2, V) (:101-112), 16-point
 *                         rule; the P1 mass solve that finishes the projection is tm_filter_apply
 *                         (kind 1) of an engine created with filter_radius = 0
 *   tm_fluid_apply        the operator itself (mode 0) / the lifting of boundary values (mode 1),
 *                         for the parity tests */
typedef struct tm_fluid_s* tm_fluid_handle;
int tm_fluid_create(int nx, int ny, double width, double height, double viscosity, double r_min, double r_max,
                    int device, tm_fluid_handle* out);
int tm_fluid_destroy(tm_fluid_handle h);
int tm_fluid_set_stream(tm_fluid_handle h, void* stream);
/* option 1: MINRES preconditioner, 0 = diagonal (default), 1 = multigrid on per-triangle Galerkin
 * matrices (one V-cycle per velocity component on M_r + K; pressure: M_p^-1 + a V-cycle on the P1
 * Laplacian int (1/r) grad.grad) -- needs cell counts with enough factors of two; options 2 / 3:
 * Chebyshev-Jacobi smoothing steps on the finest / the coarser levels (default 2 / 3); option 4:
 * warm start from the previous solve's solution (default 0), tolerance still relative to the
 * original right-hand side, dropped when its residual is not smaller than the zero guess's;
 * option 5: keep the MINRES scalars on the device (one-thread scalar steps, no synchronisation per
 * iteration; default 0) and option 6: iterations between the residual read-backs then (default 10);
 * option 7: deterministic mode -- gather kernels (one work item per output entry, fixed summation
 * order, no atomics) instead of the per-triangle scatter kernels (default 0); option 8: six MINRES
 * iterations (the period of the vector roles) captured once per solve and replayed as a CUDA graph
 * (implies option 5; default 0) */
int tm_fluid_set_option(tm_fluid_handle h, int option, double value);
int tm_fluid_set_density(tm_fluid_handle h, const double* rho, double q);
int tm_fluid_state_solve(tm_fluid_handle h, const double* boundary_velocity, double rtol, int maxit, double* up,
                         int* iters, double* relres);
int tm_fluid_objective(tm_fluid_handle h, const double* u, double* out);
int tm_fluid_sens_rhs(tm_fluid_handle h, const double* rho, const double* u, double* out);
int tm_fluid_apply(tm_fluid_handle h, const double* x, double* y, int mode);

/* the Q1 strain-energy evaluator of the reference's deep-energy back-end, float32 as
 * there (replaces StrainEnergy.calculate_objective_and_gradient / the internal part of
 * calculate_energy, reference: DEM_src/elasisity_problem.py:82-129 over
 * DEM_src/objective_calculator.py:51-143).  Device pointers; layouts are the reference's:
 * u[(ix*(ny+1)+iy)*2 + c], density / cell_energy / grad_density [iy][ix], grad_u like u.
 *   cell_energy  = det J * sum over the 2x2 Gauss points of sigma:eps        (may be NULL)
 *   grad_density = -r'(rho) * cell_energy,  r = m + rho^p (1-m)              (may be NULL)
 *   *objective   = sum r(rho) * cell_energy  (host double; internal energy = objective / 2)
 *   grad_u       = d(objective / 2)/du, what the reference obtains by autograd (may be NULL)
 * Synchronises `stream` before returning. */
int tm_dem_strain_energy(int nx, int ny, double width, double height, double lame_lambda, double lame_mu,
                         double simp_min, double penalty, const float* u, const float* density,
                         float* cell_energy, float* grad_density, float* grad_u, double* objective,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif
