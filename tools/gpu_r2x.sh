# Round 2, call X (1 GPU, the last 2.7 GPU-minutes): the relaxed fused-rz test, the small cycle-window case, racecheck on it
# (cluster tail as a W-cycle state machine with warm re-entry).
set -x
mkdir -p gpurun_out
( timeout 60 python -m pytest tests/test_gpu_parity.py -q -s -k "fused_rz or (cycle_window and cantilever-16)" 2>&1 | tail -5 ) > gpurun_out/r2x_pytest.txt; cat gpurun_out/r2x_pytest.txt
( timeout 75 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -x -k "cycle_window and cantilever-16" 2>&1 | tail -6 ) > gpurun_out/r2x_racecheck_tail_wcycle.txt; cat gpurun_out/r2x_racecheck_tail_wcycle.txt
