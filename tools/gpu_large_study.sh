# Tuning study of the V-cycle at a bandwidth-bound size (default short_cantilever N=1792, ~51 M dofs):
#   bash tools/gpu_large_study.sh name:"opts" ...      (DESIGN_ARGS overrides the workload)
set -x
mkdir -p gpurun_out
ARGS=${DESIGN_ARGS:---design designs/short_cantilever.json --N 1792 --steps 3 --warmup 8}
show='import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[1],"iter/s",round(d["value"],3),"ms",round(d["ms_per_step"],2),d["pcg"]["iterations_by_solve"],"ms/pcg-it",round(d["ms_per_step"]/d["pcg"]["iterations_per_step"],3))'
for spec in "$@"; do
  name=${spec%%:*}; opt=${spec#*:}
  timeout 200 python bench.py $ARGS --no_cpu_baseline --no_e2e --no_mixed_leg $opt > gpurun_out/large_$name.json 2> gpurun_out/large_$name.err
  tail -c 200 gpurun_out/large_$name.err; python -c "$show" gpurun_out/large_$name.json
done
