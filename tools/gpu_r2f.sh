# Round 2, call F (4 GPUs): why is the graph-replayed sharded step (139 ms) slower than the instrumented,
# un-graphed one (110 ms)?  A/B of the V-cycle graphs on sharded engines (option 116 = graphs with collectives
# inside, 110 = all graphs), peer-memory transport (default now) and NCCL.
set -x
mkdir -p gpurun_out
run() {  # tag, env TM_P2P, extra args
  TM_P2P=$2 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 \
    --master-port 2980$4 bench.py --gpus 4 --lean --no_parity --steps 5 --warmup 3 $3 \
    > gpurun_out/r2f_$1.json 2> gpurun_out/r2f_$1.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2f_$1.json").read().strip().splitlines()[-1])
    print("$1:", round(d["ms_per_step"], 2), "ms/step timed;", round(d["roofline"]["instrumented_step_ms"], 2), "ms instrumented;", d["config"]["parallelism"][:70])
except Exception as e:
    print("$1 failed", e)
PY
}
run p2p_nographs 1 "--engine_option 116=0" 1
run p2p_nographs_at_all 1 "--engine_option 110=0" 2
run nccl_nographs 0 "--engine_option 116=0" 3
run p2p_graphs 1 "" 4
grep -c "NCCL INFO" gpurun_out/r2f_p2p_graphs.err; grep "Init COMPLETE" gpurun_out/r2f_p2p_graphs.err | head -3 | cut -c1-250
