"""Drop-in alias of the reference's ``FEM_src`` package: the same import paths
(``from FEM_src.solver import FEMSolver`` ...) resolve to the B200 implementation."""
