# Round 2, call Y (1 GPU, the last GPU-minute): ncu launch list of ONE timed step of short_cantilever N=512, final code.
mkdir -p gpurun_out
TM_PROFILER_RANGE=1 timeout 48 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r2y_launches_n512_lean_steps1.csv python bench.py --lean --no_parity --design short_cantilever --N 512 --steps 1 --warmup 5 > gpurun_out/r2y_launches.log 2>&1
wc -l gpurun_out/r2y_launches_n512_lean_steps1.csv
