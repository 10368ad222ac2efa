# Round 2, call B (1 GPU): the whole GPU suite on the new library (ledger, corner tractions, ADVICE fixes,
# benchmark-size checks), the bench line as the driver runs it, smoother studies on bridge N=2048, and the ncu
# launch list of the same command.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_event_reasons.active --format=csv,noheader
( time timeout 900 python -m pytest tests -x -q -m gpu -rs ) > gpurun_out/r2b_pytest_gpu.txt 2>&1; tail -8 gpurun_out/r2b_pytest_gpu.txt
( time timeout 800 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err ) 2>&1 | tail -4
tail -c 600 gpurun_out/r2b_bench.err; cut -c1-400 gpurun_out/r2b_bench.json
for opt in "109=2" "2=2" "109=2 --engine_option 2=2" "109=4"; do
  tag=$(echo "$opt" | tr -c 'a-zA-Z0-9\n' '_')
  timeout 300 python bench.py --lean --no_parity --steps 5 --warmup 3 --engine_option $opt > gpurun_out/r2b_study_$tag.json 2> gpurun_out/r2b_study_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2b_study_$tag.json").read().strip().splitlines()[-1])
    print("study $opt:", round(d["ms_per_step"], 1), "ms/step", d["pcg"]["iterations_by_solve"], "step frac", round(d["roofline"]["step"]["frac"], 3))
except Exception as e:
    print("study $opt failed", e)
PY
done
TM_PROFILER_RANGE=1 timeout 500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r2b_launches_bench_lean_steps1.csv python bench.py --lean --no_parity --steps 1 --warmup 3 > gpurun_out/r2b_launches.log 2>&1
wc -l gpurun_out/r2b_launches_bench_lean_steps1.csv; tail -3 gpurun_out/r2b_launches.log | cut -c1-300
ls -la gpurun_out/ | tail -12
