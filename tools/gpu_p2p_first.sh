# First multi-GPU call of round 2 (the 1-GPU part also runs the opt-in fluid multigrid test)
# First multi-GPU call of round 2 (gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_p2p_first.sh'):
# the cross-GPU part of the peer-memory transport (csrc/tm_p2p.cuh) has not run on hardware yet.
#  1. loop-back self-test on one GPU (passed in round 1), as a canary;
#  2. sharded-vs-single check under NCCL and under the peer-memory transport (TM_TEST_P2P=1);
#  3. A/B of the 2-GPU weak-scaling bench line: NCCL vs TM_P2P=1.
# Every step runs under its own timeout: a poll that never completes raises after ~4 s per rank.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
nvidia-smi topo -m | head -12
timeout 120 python -m pytest tests/test_gpu_z_p2p_loopback.py -q 2>&1 | tail -3
TM_TEST_FLUID_MG=1 timeout 200 python -m pytest tests/test_gpu_z_fluid.py -q -k "multigrid or optin" 2>&1 | tail -15 | tee gpurun_out/r2_fluid_mg.txt
TM_TEST_P2P=1 timeout 700 python -m pytest tests/test_gpu_sharded.py -q -x 2>&1 | tail -15 | tee gpurun_out/r2_p2p_sharded.txt
for mode in 0 1; do
  TM_P2P=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 2960$mode bench.py --gpus 2 --steps 5 --warmup 3 --no_cpu_baseline --no_mixed_leg \
    > gpurun_out/r2_bench_2gpu_p2p$mode.json 2> gpurun_out/r2_bench_2gpu_p2p$mode.err
  tail -c 400 gpurun_out/r2_bench_2gpu_p2p$mode.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench_2gpu_p2p$mode.json").read().strip().splitlines()[-1])
print("TM_P2P=$mode", round(d["value"], 2), "iter/s normalised,", round(d["ms_per_step"], 2), "ms/step;", d["config"]["parallelism"])
PY
done
