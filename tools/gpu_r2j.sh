# Round 2, call J (1 GPU): prolongation with the weight table in shared memory (vs 67 ms / 3.4 TB/s per step in
# call I), the full GPU suite, smoke(), and the CPU reference arm exactly as the driver launches it.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_event_reasons.active --format=csv,noheader
timeout 300 python bench.py --lean --no_parity --steps 8 --warmup 4 > gpurun_out/r2j_base.json 2> gpurun_out/r2j_base.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2j_base.json").read().strip().splitlines()[-1])
    bc = d["roofline"]["by_category_one_instrumented_step"]
    print("base:", round(d["ms_per_step"], 1), "ms/step", d["pcg"]["iterations_by_solve"], "| instrumented", round(d["roofline"]["instrumented_step_ms"], 1),
          {c: (bc[c]["ms"], bc[c]["GBps"]) for c in ("level1_op", "level2_op", "restrict", "prolong", "filter")}, "step frac", d["roofline"]["step"]["frac"])
except Exception as e:
    print("base failed", e)
PY
( time timeout 900 python -m pytest tests -x -q -m gpu -rs ) > gpurun_out/r2j_pytest_gpu.txt 2>&1; tail -8 gpurun_out/r2j_pytest_gpu.txt
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -5
( time timeout 860 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2j_reference_arm.json 2> gpurun_out/r2j_reference_arm.err ) 2>&1 | tail -4
tail -c 300 gpurun_out/r2j_reference_arm.err; cut -c1-900 gpurun_out/r2j_reference_arm.json
ls -la gpurun_out/ | tail -5
