# the module name keeps the reference's spelling (FEM_src/elasisity_problem.py)
from topomax_b200.elasticity_problem import ElasticityProblem  # noqa: F401
