from topomax_b200.fem_solver import FEMSolver  # noqa: F401  (reference: FEM_src/solver.py:15)
