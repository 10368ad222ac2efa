# Round 2, call C (2 GPUs): whole GPU suite incl. the sharded checks under NCCL and under the peer-memory
# transport (first run of the IPC mapping on real peers), then the 2-GPU bench line (triangle N=4096) and a
# latency-bound mesh, both transports.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
nvidia-smi topo -m | head -8
( time timeout 1200 python -m pytest tests -x -q -m gpu -rs ) > gpurun_out/r2c_pytest_gpu.txt 2>&1; tail -12 gpurun_out/r2c_pytest_gpu.txt
for mode in 0 1; do
  TM_P2P=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 2960$mode bench.py --gpus 2 --steps 5 --warmup 3 --no_e2e --no_mixed_leg \
    > gpurun_out/r2c_bench_2gpu_p2p$mode.json 2> gpurun_out/r2c_bench_2gpu_p2p$mode.err
  grep -v "NCCL INFO" gpurun_out/r2c_bench_2gpu_p2p$mode.err | tail -c 600
  TM_P2P=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 2961$mode bench.py --gpus 2 --design short_cantilever --N 720 --lean --no_parity --steps 5 --warmup 3 \
    > gpurun_out/r2c_bench_2gpu_small_p2p$mode.json 2> gpurun_out/r2c_bench_2gpu_small_p2p$mode.err
  python - <<PY
import json
for f in ("gpurun_out/r2c_bench_2gpu_p2p$mode.json", "gpurun_out/r2c_bench_2gpu_small_p2p$mode.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("TM_P2P=$mode", d["config"]["workload"][:40], round(d["ms_per_step"], 2), "ms/step;", d["config"]["parallelism"][:60],
              "| parity", (d.get("parity") or {}).get("ok"), (d.get("parity") or {}).get("relative_residual"),
              "| 1gpu", (d.get("single_gpu_comparison") or {}).get("objective_trace_max_rel_diff"),
              (d.get("single_gpu_comparison") or {}).get("strong_scaling_speedup"))
        print("   phases", d["roofline"]["phases_one_instrumented_step_ms"])
    except Exception as e:
        print("failed", f, e)
PY
done
ls -la gpurun_out/ | tail -8
