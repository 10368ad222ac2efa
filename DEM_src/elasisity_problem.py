from topomax_b200.dem_energy import StrainEnergy  # noqa: F401  (reference: DEM_src/elasisity_problem.py:64-129)
