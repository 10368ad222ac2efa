from topomax_b200.dem_energy import Mesh  # noqa: F401  (reference: DEM_src/utils.py:11-34)
