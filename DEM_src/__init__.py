"""Import-path compatibility with the reference's ``DEM_src`` package for the one piece of the
deep-energy back-end that runs on the CUDA library: the Q1 strain-energy evaluator."""
