# warm-start study: PCG iterations and step time with (a) the guarded warm start (default),
# (b) the unguarded one, (c) zero initial guesses; later optimisation iterations too
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
show='import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[1],round(d["value"],2),round(d["ms_per_step"],2),d["e2e"]["value"],d["pcg"]["iterations_by_solve"],d["pcg"]["warm_starts_kept"])'
for spec in guard: noguard:--engine_option=125=0 zero:--no_warm_start; do
  name=${spec%%:*}; opt=${spec#*:}
  timeout 200 python bench.py --steps 5 --warmup 3 --no_cpu_baseline --no_e2e $opt > gpurun_out/ws_$name.json 2> gpurun_out/ws_$name.err
  tail -c 300 gpurun_out/ws_$name.err; python -c "$show" gpurun_out/ws_$name.json
  timeout 200 python bench.py --steps 5 --warmup 25 --no_cpu_baseline --no_e2e $opt > gpurun_out/ws25_$name.json 2> gpurun_out/ws25_$name.err
  tail -c 300 gpurun_out/ws25_$name.err; python -c "$show" gpurun_out/ws25_$name.json
done
timeout 300 python bench.py --design bridge --N 2048 --steps 3 --warmup 3 --no_cpu_baseline > gpurun_out/r1i_bench_bridge2048.json 2> gpurun_out/r1i_bridge.err
tail -c 300 gpurun_out/r1i_bridge.err; python -c "$show" gpurun_out/r1i_bench_bridge2048.json
