from topomax_b200.fluid_problem import BoundaryFlows, FluidProblem  # noqa: F401  (reference: FEM_src/fluid_problem.py)
