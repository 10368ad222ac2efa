"""topomax_b200: B200-native FEM elasticity inner loop of topomax (drop-in for the reference's
``FEM_src`` Solver/Problem/filter interface).  CUDA kernels live in ``csrc/`` behind the C ABI of
``include/topomax_b200.h``; importing the package does not touch the GPU, constructing a solver
or an engine does -- and fails loudly without one."""

__all__ = ["FEMSolver", "ElasticityProblem", "HelmholtzFilter", "Engine", "parse_design"]


def __getattr__(name):
    if name == "FEMSolver":
        from .fem_solver import FEMSolver
        return FEMSolver
    if name == "ElasticityProblem":
        from .elasticity_problem import ElasticityProblem
        return ElasticityProblem
    if name == "HelmholtzFilter":
        from .filter import HelmholtzFilter
        return HelmholtzFilter
    if name == "Engine":
        from .engine import Engine
        return Engine
    if name == "parse_design":
        from .designs.design_parser import parse_design
        return parse_design
    raise AttributeError(name)
