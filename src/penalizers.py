from topomax_b200.penalizers import ElasticPenalizer, Penalizer  # noqa: F401
