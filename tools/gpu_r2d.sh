# Round 2, call D (1 GPU): tiled restriction (bit-identity test + A/B), the full GPU suite, ncu --set full of
# the transfer kernels at bridge N=2048, the bench line as the driver runs it, the latency-bound config's
# per-category breakdown, and the fluid row's measurement.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_event_reasons.active --format=csv,noheader
( time timeout 900 python -m pytest tests -x -q -m gpu -rs ) > gpurun_out/r2d_pytest_gpu.txt 2>&1; tail -8 gpurun_out/r2d_pytest_gpu.txt
for opt in "128=0" "128=1"; do
  tag=$(echo "$opt" | tr -c 'a-zA-Z0-9\n' '_')
  timeout 300 python bench.py --lean --no_parity --steps 5 --warmup 3 --engine_option $opt > gpurun_out/r2d_ab_$tag.json 2> gpurun_out/r2d_ab_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2d_ab_$tag.json").read().strip().splitlines()[-1])
    bc = d["roofline"]["by_category_one_instrumented_step"]
    print("A/B $opt:", round(d["ms_per_step"], 1), "ms/step", d["pcg"]["iterations_by_solve"], "restrict", bc.get("restrict"), "prolong", bc.get("prolong"))
except Exception as e:
    print("A/B $opt failed", e)
PY
done
for spec in restrict:mg_restrict_tiled_kernelId prolong:mg_prolong_add_kernelId lvl2cheb:elast_apply_kernelIdLb1ELi3; do
  name=${spec%%:*}; rx=${spec#*:}
  skip=0; [ $name = lvl2cheb ] && skip=6
  TM_PROFILER_RANGE=1 timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none \
    --kernel-name-base mangled -k regex:$rx -s $skip -c 1 -f -o gpurun_out/r2d_$name \
    python bench.py --lean --no_parity --steps 1 --warmup 3 > gpurun_out/r2d_$name.log 2>&1
  tail -2 gpurun_out/r2d_$name.log | cut -c1-200
done
( time timeout 800 python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err ) 2>&1 | tail -4
tail -c 400 gpurun_out/r2d_bench.err; cut -c1-300 gpurun_out/r2d_bench.json
timeout 200 python bench.py --design short_cantilever --N 512 --lean --steps 20 --warmup 5 > gpurun_out/r2d_bench_n512_lean.json 2> gpurun_out/r2d_bench_n512_lean.err
cut -c1-200 gpurun_out/r2d_bench_n512_lean.json
for args in "--N 128" "--N 128 --preconditioner multigrid --graph" "--N 256" "--N 256 --preconditioner multigrid --graph" "--N 256 --preconditioner multigrid --graph --deterministic"; do
  timeout 300 python tools/fluid_bench.py $args 2>&1 | tail -1 | cut -c1-600 | tee -a gpurun_out/r2d_fluid_bench.txt
done
ls -la gpurun_out/ | tail -12
