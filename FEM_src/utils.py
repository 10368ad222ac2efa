from topomax_b200.fem_solver import load_function, save_function  # noqa: F401
from topomax_b200.sampling import mesh_to_N, mesh_to_domain_size, sample_function  # noqa: F401
