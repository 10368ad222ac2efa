# Round 2, call N (8 GPUs): the final code on the north-star configuration as the driver will run it
# (--steps 20 --warmup 5), with the full-strip CPU-operator check when the host has the memory; then 4 GPUs.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader | head -2; free -g | head -2; nproc
( time timeout 870 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29911 \
    bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2n_bench_8gpu.json 2> gpurun_out/r2n_bench_8gpu.err ) 2>&1 | tail -3
grep -v "NCCL INFO" gpurun_out/r2n_bench_8gpu.err | tail -c 1200
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2n_bench_8gpu.json").read().strip().splitlines()[-1])
    print("8 GPUs:", d["config"]["workload"], round(d["ms_per_step"], 1), "ms/step, value", round(d["value"], 3), "e2e", d["e2e"]["value"],
          "| its", d["pcg"]["iterations_by_solve"], "| step frac", round(d["roofline"]["step"]["frac"], 3), "| dominant", d["roofline"]["category"], round(d["roofline"]["frac"], 3))
    print("   parity", d["parity"])
    print("   phases", d["roofline"]["phases_one_instrumented_step_ms"])
except Exception as e:
    print("8-GPU line failed", e)
PY
sleep 5
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29912 \
    bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2n_bench_4gpu.json 2> gpurun_out/r2n_bench_4gpu.err ) 2>&1 | tail -3
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2n_bench_4gpu.json").read().strip().splitlines()[-1])
    print("4 GPUs:", d["config"]["workload"], round(d["ms_per_step"], 1), "ms/step, value", round(d["value"], 3), "e2e", d["e2e"]["value"],
          "| parity", d["parity"]["ok"], d["parity"]["relative_residual"], "| 1gpu", d["single_gpu_comparison"])
except Exception as e:
    print("4-GPU line failed", e)
PY
