from topomax_b200.designs.definitions import *  # noqa: F401,F403
from topomax_b200.designs.definitions import to_2_tuple  # noqa: F401
