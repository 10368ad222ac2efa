# IterationData / SolverResult are pickled under this module path by reference and port alike
from topomax_b200.utils import (  # noqa: F401
    IterationData, SolverResult, Timer, get_solver_data, prettify_seconds, smart_brentq,
)
