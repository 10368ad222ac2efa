from topomax_b200.problem import Problem  # noqa: F401
