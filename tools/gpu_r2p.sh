# Round 2, call P (1 GPU): tiled prolongation (option 132): A/B on bridge N=2048 before anything else, then the
# whole GPU suite (incl. its bit-identity test) and racecheck on it.
set -x
mkdir -p gpurun_out
for opt in "132=0" "132=1" "132=0" "132=1"; do
  tag=$(echo "$opt" | tr -c 'a-zA-Z0-9\n' '_')_$RANDOM
  timeout 300 python bench.py --lean --no_parity --steps 5 --warmup 3 --engine_option $opt > gpurun_out/r2p_ab_$tag.json 2> gpurun_out/r2p_ab_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2p_ab_$tag.json").read().strip().splitlines()[-1])
    bc = d["roofline"]["by_category_one_instrumented_step"]
    print("bridge $opt:", round(d["ms_per_step"], 1), "ms/step", d["pcg"]["iterations_by_solve"], "| instrumented", round(d["roofline"]["instrumented_step_ms"], 1),
          {c: (bc[c]["ms"], bc[c]["GBps"]) for c in ("prolong", "restrict", "level1_op") if c in bc}, "step frac", round(d["roofline"]["step"]["frac"], 3))
except Exception as e:
    print("bridge $opt failed", e); print(open("gpurun_out/r2p_ab_$tag.err").read()[-800:])
PY
done
( time timeout 900 python -m pytest tests -x -q -m gpu -rs ) > gpurun_out/r2p_pytest_gpu.txt 2>&1; tail -8 gpurun_out/r2p_pytest_gpu.txt
( timeout 300 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -x -k "tiled_prolongation and (150-37 or 13-7 or 256-64)" 2>&1 | tail -6 ) > gpurun_out/r2p_racecheck.txt; tail -3 gpurun_out/r2p_racecheck.txt
