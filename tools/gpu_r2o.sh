# Round 2, call O (1 GPU): the final code, exactly the driver's sequence: smoke(), the bench line
# (--steps 20 --warmup 5), the ncu launch list of the same command (one timed step).
set -x
mkdir -p gpurun_out
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -4
( time timeout 800 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err ) 2>&1 | tail -4
tail -c 300 gpurun_out/r2o_bench.err; cut -c1-300 gpurun_out/r2o_bench.json
TM_PROFILER_RANGE=1 timeout 500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r2o_launches_bench_lean_steps1.csv python bench.py --lean --no_parity --steps 1 --warmup 5 > gpurun_out/r2o_launches.log 2>&1
wc -l gpurun_out/r2o_launches_bench_lean_steps1.csv
