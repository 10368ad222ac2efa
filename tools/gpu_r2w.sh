# Round 2, call W (1 GPU): the final code exactly as the driver runs it: smoke(), the bench line (--steps 20 --warmup 5),
# A/B of the attainable-accuracy stop (option 137) with the CPU-operator residual of both, the whole GPU suite, the
# ncu launch list of one timed step.
set -x
mkdir -p gpurun_out
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -4
( time timeout 800 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err ) 2>&1 | tail -4
tail -c 300 gpurun_out/r2w_bench.err; cut -c1-300 gpurun_out/r2w_bench.json
for opt in "137=0" "137=0.5"; do
  tag=$(echo "$opt" | tr -c 'a-zA-Z0-9\n' '_')
  timeout 300 python bench.py --lean --steps 5 --warmup 5 --engine_option $opt > gpurun_out/r2w_ab_floor_$tag.json 2> gpurun_out/r2w_ab_floor_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2w_ab_floor_$tag.json").read().strip().splitlines()[-1])
    print("bridge $opt:", round(d["ms_per_step"], 1), "ms/step", d["pcg"]["iterations_by_solve"], "rtol used", d["pcg"].get("rtol_used_last_solve"),
          "floor est", d["pcg"].get("fp_floor_estimate_last_solve"), "| CPU operator: residual", d["parity"]["relative_residual"], "floor", d["parity"]["fp64_floor"],
          "compliance diff", d["parity"]["compliance_rel_diff"], d["parity"]["ok"])
except Exception as e:
    print("bridge $opt failed", e); print(open("gpurun_out/r2w_ab_floor_$tag.err").read()[-800:])
PY
done
( time timeout 900 python -m pytest tests -q -m gpu -rs ) > gpurun_out/r2w_pytest_gpu.txt 2>&1; tail -8 gpurun_out/r2w_pytest_gpu.txt
TM_PROFILER_RANGE=1 timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r2w_launches_bench_lean_steps1.csv python bench.py --lean --no_parity --steps 1 --warmup 5 > gpurun_out/r2w_launches.log 2>&1
wc -l gpurun_out/r2w_launches_bench_lean_steps1.csv
