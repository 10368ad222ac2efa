# A/B study of engine options on the default bench workload: bash tools/gpu_ab.sh name:opts ...
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
show='import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[1],round(d["value"],2),round(d["ms_per_step"],2),d["e2e"]["value"],d["pcg"]["iterations_by_solve"],round(d["roofline"]["avg_launch_ms"]*1e3,1))'
for spec in "$@"; do
  name=${spec%%:*}; opt=${spec#*:}
  timeout 200 python bench.py --steps 5 --warmup 3 --no_cpu_baseline $opt > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  tail -c 300 gpurun_out/ab_$name.err; python -c "$show" gpurun_out/ab_$name.json
done
