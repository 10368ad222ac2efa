# Round 2, call K (1 GPU): sanitizers on the kernels written this round (tiled restriction with cp.async,
# prolongation with the shared-memory table, device Newton update), the new fluid tests, the whole GPU suite.
set -x
mkdir -p gpurun_out
K='tiled_restriction and 150-37 or tiled_restriction and 13-7 or multigrid_galerkin_and_transfer_adjoint or device_resident_newton or state_solve_matches_direct_solver and multigrid and triangle'
( timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -x -k "$K" 2>&1 | tail -12 ) > gpurun_out/r2k_memcheck.txt; tail -4 gpurun_out/r2k_memcheck.txt
( timeout 500 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -x -k "$K" 2>&1 | tail -12 ) > gpurun_out/r2k_racecheck.txt; tail -4 gpurun_out/r2k_racecheck.txt
( time timeout 900 python -m pytest tests -x -q -m gpu -rs ) > gpurun_out/r2k_pytest_gpu.txt 2>&1; tail -8 gpurun_out/r2k_pytest_gpu.txt
ls -la gpurun_out/ | tail -4
