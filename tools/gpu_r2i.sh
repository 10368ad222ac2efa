# Round 2, call I (1 GPU): prolongation with 4 rows per thread (vs 53 ms / 3.4 TB/s in call G), the extrapolated
# warm start study, ncu of the level-0 prolongation, the full GPU suite, and the bench line as the driver runs it.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_event_reasons.active --format=csv,noheader
for extra in "" "--extrapolate"; do
  tag=base; [ -n "$extra" ] && tag=extrapolate
  timeout 300 python bench.py --lean --no_parity --steps 8 --warmup 4 $extra > gpurun_out/r2i_$tag.json 2> gpurun_out/r2i_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2i_$tag.json").read().strip().splitlines()[-1])
    bc = d["roofline"]["by_category_one_instrumented_step"]
    print("$tag:", round(d["ms_per_step"], 1), "ms/step", d["pcg"]["iterations_by_solve"], "kept", d["pcg"]["warm_starts_kept"], "| instrumented", round(d["roofline"]["instrumented_step_ms"], 1),
          {c: (bc[c]["ms"], bc[c]["GBps"]) for c in ("level1_op", "level2_op", "restrict", "prolong", "filter")})
except Exception as e:
    print("$tag failed", e)
PY
done
TM_PROFILER_RANGE=1 timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none \
  --kernel-name-base mangled -k regex:^_ZN3tmx21mg_prolong_add -s 7 -c 1 -f -o gpurun_out/r2i_prolong \
  python bench.py --lean --no_parity --steps 1 --warmup 3 > gpurun_out/r2i_prolong.log 2>&1
( time timeout 900 python -m pytest tests -x -q -m gpu -rs ) > gpurun_out/r2i_pytest_gpu.txt 2>&1; tail -8 gpurun_out/r2i_pytest_gpu.txt
sleep 3
( time timeout 800 python bench.py --steps 20 --warmup 5 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err ) 2>&1 | tail -4
tail -c 300 gpurun_out/r2i_bench.err; cut -c1-300 gpurun_out/r2i_bench.json
TM_PROFILER_RANGE=1 timeout 500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r2i_launches_bench_lean_steps1.csv python bench.py --lean --no_parity --steps 1 --warmup 5 > gpurun_out/r2i_launches.log 2>&1
wc -l gpurun_out/r2i_launches_bench_lean_steps1.csv
ls -la gpurun_out/ | tail -6
