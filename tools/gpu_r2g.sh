# Round 2, call G (1 GPU): A/B of the wave-aware strip heights (option 129) and of the cp.async tiled restriction
# on bridge N=2048 BEFORE anything else runs on the box, ncu of the prolongation, then the whole GPU suite.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_event_reasons.active --format=csv,noheader
for opt in "129=0" "129=1" "129=0" "129=1"; do
  tag=$(echo "$opt" | tr -c 'a-zA-Z0-9\n' '_')_$RANDOM
  timeout 300 python bench.py --lean --no_parity --steps 5 --warmup 3 --engine_option $opt > gpurun_out/r2g_ab_$tag.json 2> gpurun_out/r2g_ab_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2g_ab_$tag.json").read().strip().splitlines()[-1])
    bc = d["roofline"]["by_category_one_instrumented_step"]
    print("A/B $opt:", round(d["ms_per_step"], 1), "ms/step", d["pcg"]["iterations_by_solve"], "| instrumented", round(d["roofline"]["instrumented_step_ms"], 1),
          {c: (bc[c]["ms"], bc[c]["GBps"]) for c in ("level1_op", "level2_op", "level3_op", "level4_op", "restrict", "prolong")})
except Exception as e:
    print("A/B $opt failed", e)
PY
done
TM_PROFILER_RANGE=1 timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none \
  --kernel-name-base mangled -k regex:^_ZN3tmx21mg_prolong_add -c 1 -f -o gpurun_out/r2g_prolong \
  python bench.py --lean --no_parity --steps 1 --warmup 3 > gpurun_out/r2g_prolong.log 2>&1
TM_PROFILER_RANGE=1 timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none \
  --kernel-name-base mangled -k regex:mg_restrict_tiled -c 1 -f -o gpurun_out/r2g_restrict \
  python bench.py --lean --no_parity --steps 1 --warmup 3 > gpurun_out/r2g_restrict.log 2>&1
( time timeout 900 python -m pytest tests -x -q -m gpu -rs ) > gpurun_out/r2g_pytest_gpu.txt 2>&1; tail -8 gpurun_out/r2g_pytest_gpu.txt
ls -la gpurun_out/ | tail -8
