from topomax_b200.fem_solver import load_function, save_function  # noqa: F401
