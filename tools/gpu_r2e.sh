# Round 2, call E (4 GPUs): sharded-vs-single checks at 2 and 4 ranks under both transports, then the 4-GPU bench
# line (triangle N=4096: BASELINE config 4) under NCCL and under the peer-memory transport.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
( time timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -rs ) > gpurun_out/r2e_pytest_sharded.txt 2>&1; tail -8 gpurun_out/r2e_pytest_sharded.txt
for mode in 0 1; do
  TM_P2P=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 \
    --master-port 2970$mode bench.py --gpus 4 --steps 5 --warmup 3 --no_e2e --no_mixed_leg \
    > gpurun_out/r2e_bench_4gpu_p2p$mode.json 2> gpurun_out/r2e_bench_4gpu_p2p$mode.err
  grep -v "NCCL INFO" gpurun_out/r2e_bench_4gpu_p2p$mode.err | tail -c 400
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2e_bench_4gpu_p2p$mode.json").read().strip().splitlines()[-1])
    print("TM_P2P=$mode", d["config"]["workload"][:40], round(d["ms_per_step"], 2), "ms/step;", d["config"]["parallelism"][:60],
          "| parity", (d.get("parity") or {}).get("ok"), (d.get("parity") or {}).get("relative_residual"),
          "| 1gpu", (d.get("single_gpu_comparison") or {}).get("objective_trace_max_rel_diff"),
          (d.get("single_gpu_comparison") or {}).get("strong_scaling_speedup"))
    print("   phases", d["roofline"]["phases_one_instrumented_step_ms"])
except Exception as e:
    print("failed", e)
PY
done
grep -c "NCCL INFO" gpurun_out/r2e_bench_4gpu_p2p0.err; grep "NCCL INFO.*nranks\|NCCL INFO.*Init COMPLETE" gpurun_out/r2e_bench_4gpu_p2p0.err | head -4
ls -la gpurun_out/ | tail -6
