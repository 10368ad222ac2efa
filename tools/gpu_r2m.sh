# Round 2, call M (2 GPUs): the final code on two real peers -- bench first (triangle N=4096, both transports),
# then the sharded-vs-single checks (EP_CHEB0 exchanges b's halo rows on the sharded levels).
set -x
mkdir -p gpurun_out
for mode in 1 0; do
  TM_P2P=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 2990$mode bench.py --gpus 2 --steps 5 --warmup 3 --no_mixed_leg \
    > gpurun_out/r2m_bench_2gpu_p2p$mode.json 2> gpurun_out/r2m_bench_2gpu_p2p$mode.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2m_bench_2gpu_p2p$mode.json").read().strip().splitlines()[-1])
    print("TM_P2P=$mode", d["config"]["workload"][:40], round(d["ms_per_step"], 2), "ms/step; e2e", d["e2e"]["value"], d["config"]["parallelism"][:50],
          "| parity", d["parity"]["ok"], d["parity"]["relative_residual"], "| 1gpu", d["single_gpu_comparison"]["objective_trace_max_rel_diff"],
          d["single_gpu_comparison"]["strong_scaling_speedup"], d["single_gpu_comparison"]["ms_per_step_one_gpu"])
except Exception as e:
    print("failed", e)
PY
  sleep 3
done
( time timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -rs ) > gpurun_out/r2m_pytest_sharded.txt 2>&1; tail -6 gpurun_out/r2m_pytest_sharded.txt
