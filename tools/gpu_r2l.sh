# Round 2, call L (1 GPU): EP_CHEB0 (first two smoothing steps of the coarse levels in one operator pass) and
# three resident blocks per SM for the stored-moment kernels: A/B on bridge N=2048 and short_cantilever N=512,
# then the whole GPU suite (incl. the fused-vs-separate test) and the fluid defaults.
set -x
mkdir -p gpurun_out
for opt in "131=0" "131=1" "131=0" "131=1"; do
  tag=$(echo "$opt" | tr -c 'a-zA-Z0-9\n' '_')_$RANDOM
  timeout 300 python bench.py --lean --no_parity --steps 5 --warmup 3 --engine_option $opt > gpurun_out/r2l_ab_$tag.json 2> gpurun_out/r2l_ab_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2l_ab_$tag.json").read().strip().splitlines()[-1])
    bc = d["roofline"]["by_category_one_instrumented_step"]
    print("bridge $opt:", round(d["ms_per_step"], 1), "ms/step", d["pcg"]["iterations_by_solve"], "| instrumented", round(d["roofline"]["instrumented_step_ms"], 1),
          {c: (bc[c]["ms"], bc[c]["GBps"]) for c in ("level1_op", "level2_op", "cheb_first") if c in bc}, "step frac", round(d["roofline"]["step"]["frac"], 3))
except Exception as e:
    print("bridge $opt failed", e)
PY
done
for opt in "131=0" "131=1" "131=0" "131=1"; do
  tag=$(echo "$opt" | tr -c 'a-zA-Z0-9\n' '_')_$RANDOM
  timeout 200 python bench.py --design short_cantilever --N 512 --lean --no_parity --steps 20 --warmup 5 --engine_option $opt > gpurun_out/r2l_n512_$tag.json 2> gpurun_out/r2l_n512_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2l_n512_$tag.json").read().strip().splitlines()[-1])
    print("N=512 $opt:", round(d["ms_per_step"], 2), "ms/step =", round(d["value"], 2), "iter/s, its/step", d["pcg"]["iterations_per_step"], "launches", d["gpu_launches"])
except Exception as e:
    print("N=512 $opt failed", e)
PY
done
( time timeout 900 python -m pytest tests -x -q -m gpu -rs ) > gpurun_out/r2l_pytest_gpu.txt 2>&1; tail -8 gpurun_out/r2l_pytest_gpu.txt
ls -la gpurun_out/ | tail -4
