# Round 2, call H (8 GPUs): the north-star configuration FIRST on the fresh box -- cantilever N=16384
# (49152 x 16384 cells, 6.44e9 displacement dofs) strip-sharded over 8 B200 -- with its self-check; then the
# 4-GPU line (triangle N=4096) with the single-GPU comparison; then the sharded-vs-single checks at 8 ranks.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29901 \
    bench.py --gpus 8 --steps 5 --warmup 3 --no_mixed_leg > gpurun_out/r2h_bench_8gpu.json 2> gpurun_out/r2h_bench_8gpu.err ) 2>&1 | tail -3
grep -v "NCCL INFO" gpurun_out/r2h_bench_8gpu.err | tail -c 1500
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2h_bench_8gpu.json").read().strip().splitlines()[-1])
    print("8 GPUs:", d["config"]["workload"], round(d["ms_per_step"], 1), "ms/step, value", round(d["value"], 3), "e2e", d["e2e"]["value"],
          "| its", d["pcg"]["iterations_by_solve"], "| step frac", round(d["roofline"]["step"]["frac"], 3), "| dominant", d["roofline"]["category"], round(d["roofline"]["frac"], 3))
    print("   parity", d["parity"])
    print("   phases", d["roofline"]["phases_one_instrumented_step_ms"])
except Exception as e:
    print("8-GPU line failed", e)
PY
nvidia-smi --query-gpu=index,memory.used --format=csv,noheader | head -3
sleep 5
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29902 \
    bench.py --gpus 4 --steps 5 --warmup 3 --no_mixed_leg > gpurun_out/r2h_bench_4gpu.json 2> gpurun_out/r2h_bench_4gpu.err ) 2>&1 | tail -3
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2h_bench_4gpu.json").read().strip().splitlines()[-1])
    print("4 GPUs:", d["config"]["workload"], round(d["ms_per_step"], 1), "ms/step, value", round(d["value"], 3), "e2e", d["e2e"]["value"],
          "| parity", d["parity"]["ok"], d["parity"]["relative_residual"], "| 1gpu", d["single_gpu_comparison"])
except Exception as e:
    print("4-GPU line failed", e)
PY
sleep 5
( time timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -rs -k "8" ) > gpurun_out/r2h_pytest_sharded8.txt 2>&1; tail -6 gpurun_out/r2h_pytest_sharded8.txt
ls -la gpurun_out/ | tail -5
