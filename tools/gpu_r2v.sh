# Round 2, call V (2 GPUs): the final defaults on two real peers: bench (triangle N=4096, peer-memory transport,
# objective trace against a single-GPU run of the same mesh), then the sharded-vs-single checks at world 2.
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 29901 bench.py --gpus 2 --steps 5 --warmup 3 --no_mixed_leg \
    > gpurun_out/r2v_bench_2gpu.json 2> gpurun_out/r2v_bench_2gpu.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2v_bench_2gpu.json").read().strip().splitlines()[-1])
    print(d["config"]["workload"][:40], round(d["ms_per_step"], 2), "ms/step; e2e", d["e2e"]["value"], d["config"]["parallelism"][:50],
          "| parity", d["parity"]["ok"], d["parity"]["relative_residual"], d["parity"].get("fp64_floor"), "| 1gpu", d["single_gpu_comparison"]["objective_trace_max_rel_diff"],
          d["single_gpu_comparison"]["strong_scaling_speedup"], d["single_gpu_comparison"]["ms_per_step_one_gpu"], "| pcg", d["pcg"])
except Exception as e:
    print("failed", e)
PY
tail -c 600 gpurun_out/r2v_bench_2gpu.err
( time timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q -rs ) > gpurun_out/r2v_pytest_sharded.txt 2>&1; tail -6 gpurun_out/r2v_pytest_sharded.txt
( time timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -s -k "attainable or cycle_window" ) > gpurun_out/r2v_pytest_new.txt 2>&1; tail -8 gpurun_out/r2v_pytest_new.txt
