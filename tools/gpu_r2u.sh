# Round 2, call U (1 GPU): where the N=512 step goes (host-visible phases, instrumented categories), the rest of the
# GPU suite, sanitizers on the W-capable tail kernel.
set -x
mkdir -p gpurun_out
timeout 300 python tools/step_phases.py short_cantilever 512 5 20 > gpurun_out/r2u_step_phases_n512.json 2> gpurun_out/r2u_step_phases_n512.err; cat gpurun_out/r2u_step_phases_n512.json; tail -3 gpurun_out/r2u_step_phases_n512.err
timeout 300 python tools/step_phases.py short_cantilever 512 5 20 135=1 136=0 > gpurun_out/r2u_step_phases_n512_vcycle.json 2>> gpurun_out/r2u_step_phases_n512.err; cat gpurun_out/r2u_step_phases_n512_vcycle.json
timeout 300 python bench.py --lean --no_parity --design short_cantilever --N 512 --steps 20 --warmup 5 > gpurun_out/r2u_bench_n512_lean.json 2> gpurun_out/r2u_bench_n512_lean.err; cut -c1-300 gpurun_out/r2u_bench_n512_lean.json
timeout 300 python tools/step_phases.py bridge 2048 5 10 > gpurun_out/r2u_step_phases_bridge2048.json 2> gpurun_out/r2u_step_phases_bridge.err; cat gpurun_out/r2u_step_phases_bridge2048.json
( time timeout 900 python -m pytest tests -q -m gpu -rs ) > gpurun_out/r2u_pytest_gpu.txt 2>&1; tail -8 gpurun_out/r2u_pytest_gpu.txt
( timeout 400 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -x -k "cluster_tail and (bridge or cantilever-48) and (2 or 3)" 2>&1 | tail -6 ) > gpurun_out/r2u_racecheck_tail.txt; tail -3 gpurun_out/r2u_racecheck_tail.txt
( timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -x -k "cycle_window or (cluster_tail and 35)" 2>&1 | tail -6 ) > gpurun_out/r2u_memcheck_cycle.txt; tail -3 gpurun_out/r2u_memcheck_cycle.txt
