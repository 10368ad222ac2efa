from topomax_b200.filter import AssembledP1Form, HelmholtzFilter  # noqa: F401
