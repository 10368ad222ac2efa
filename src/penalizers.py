from topomax_b200.penalizers import ElasticPenalizer, Penalizer  # noqa: F401
from topomax_b200.fluid_problem import FluidPenalizer  # noqa: F401,E402  (reference: src/penalizers.py:49-68)
