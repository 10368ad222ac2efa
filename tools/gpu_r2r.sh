# Round 2, call R (1 GPU): W-capable cluster tail (per-level gamma, warm re-entry): parity tests, then windows
# that now reach into the tail.
set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -s -k "cycle_window or cluster_tail" ) > gpurun_out/r2r_pytest_cycle.txt 2>&1; tail -25 gpurun_out/r2r_pytest_cycle.txt
timeout 600 python tools/cycle_study.py short_cantilever 512 25 "133=4,134=4,135=2" "133=5,134=5,135=2" "133=6,134=6,135=2" "133=5,134=6,135=2" "133=6,134=7,135=2" "133=5,134=7,135=2" \
   "133=4,134=5,135=2" "133=4,134=6,135=2" "133=5,135=2" "133=6,135=2" "133=4,135=2" "133=5,134=6,135=3" "133=6,134=6,135=3" \
   > gpurun_out/r2r_cycle_study_n512.jsonl 2> gpurun_out/r2r_cycle_study_n512.err; cat gpurun_out/r2r_cycle_study_n512.jsonl | cut -c1-300; tail -3 gpurun_out/r2r_cycle_study_n512.err
timeout 900 python tools/cycle_study.py bridge 2048 25 "133=5,134=7,135=2" "133=5,134=8,135=2" "133=5,134=9,135=2" "133=6,134=9,135=2" "133=7,134=9,135=2" "133=8,134=9,135=2" \
   "133=5,135=2" "133=6,134=8,135=3" "133=6,134=9,135=3" "133=7,134=9,135=3" \
   > gpurun_out/r2r_cycle_study_bridge2048.jsonl 2> gpurun_out/r2r_cycle_study_bridge2048.err; cat gpurun_out/r2r_cycle_study_bridge2048.jsonl | cut -c1-300; tail -3 gpurun_out/r2r_cycle_study_bridge2048.err
