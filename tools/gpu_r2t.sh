# Round 2, call T (1 GPU): final defaults (automatic W window, coarsest level <= 4 cells, two smoothing steps on
# levels 1-2 of large meshes, residual read-backs from the expected iteration count): per-level degree study, the
# whole GPU suite, the bench line.
set -x
mkdir -p gpurun_out
timeout 600 python tools/cycle_study.py short_cantilever 512 25 "136=0" "136=1" "136=2" "136=3" "136=4" "3=1" \
   > gpurun_out/r2t_cycle_study_n512.jsonl 2> gpurun_out/r2t_cycle_study_n512.err; cat gpurun_out/r2t_cycle_study_n512.jsonl | cut -c1-300; tail -3 gpurun_out/r2t_cycle_study_n512.err
timeout 900 python tools/cycle_study.py bridge 2048 25 "136=0" "136=1" "136=2" "136=3" "136=5" "109=2" \
   > gpurun_out/r2t_cycle_study_bridge2048.jsonl 2> gpurun_out/r2t_cycle_study_bridge2048.err; cat gpurun_out/r2t_cycle_study_bridge2048.jsonl | cut -c1-300; tail -3 gpurun_out/r2t_cycle_study_bridge2048.err
( time timeout 900 python -m pytest tests -x -q -m gpu -rs ) > gpurun_out/r2t_pytest_gpu.txt 2>&1; tail -8 gpurun_out/r2t_pytest_gpu.txt
( time timeout 800 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err ) 2>&1 | tail -4
tail -c 300 gpurun_out/r2t_bench.err; cut -c1-400 gpurun_out/r2t_bench.json
