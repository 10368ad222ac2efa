from topomax_b200.solver import Solver, expit, expit_diff, logit  # noqa: F401
