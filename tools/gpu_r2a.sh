# Round 2, call A (1 GPU, code of round 1): baseline of the bridge N=2048 step, its ncu launch list,
# ncu --set full of the p.Ap and level-1 operator kernels there, the fluid opt-ins, sanitizers.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_event_reasons.active --format=csv,noheader
B="--design bridge --N 2048 --no_cpu_baseline --no_mixed_leg"
timeout 500 python bench.py $B --steps 5 --warmup 3 > gpurun_out/r2a_bench_bridge2048.json 2> gpurun_out/r2a_bench_bridge2048.err
tail -c 400 gpurun_out/r2a_bench_bridge2048.err; cut -c1-600 gpurun_out/r2a_bench_bridge2048.json
TM_PROFILER_RANGE=1 timeout 500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r2a_launches_bridge2048.csv python bench.py $B --steps 1 --warmup 3 --no_e2e > gpurun_out/r2a_launches.log 2>&1
wc -l gpurun_out/r2a_launches_bridge2048.csv
for spec in dot:elast_apply_kernelIdLb0ELi1 lvl1cheb:elast_apply_kernelIdLb1ELi3; do
  name=${spec%%:*}; rx=${spec#*:}
  TM_PROFILER_RANGE=1 timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none \
    --kernel-name-base mangled -k regex:$rx -c 1 -f -o gpurun_out/r2a_$name \
    python bench.py $B --steps 1 --warmup 3 --no_e2e > gpurun_out/r2a_$name.log 2>&1
  tail -2 gpurun_out/r2a_$name.log | cut -c1-200
done
( TM_TEST_FLUID_MG=1 timeout 400 python -m pytest tests/test_gpu_z_fluid.py -q 2>&1 | tail -40 ) > gpurun_out/r2a_fluid_optins.txt; tail -5 gpurun_out/r2a_fluid_optins.txt
K='cluster_tail_equals_launch_per_phase_vcycle and cantilever-48 or temporal_blocking_matches_oracle and 40-24 or filter_matches_oracle_and_identity'
( timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_gpu_z_p2p_loopback.py -q -x -k "$K or loopback and 2-1000" 2>&1 | tail -40 ) > gpurun_out/r2a_memcheck.txt; tail -6 gpurun_out/r2a_memcheck.txt
( timeout 300 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_gpu_z_p2p_loopback.py -q -x -k "$K or loopback and 2-1000" 2>&1 | tail -40 ) > gpurun_out/r2a_racecheck.txt; tail -6 gpurun_out/r2a_racecheck.txt
ls -la gpurun_out/ | tail -20
