# Round 2, call Q (1 GPU): multigrid cycle windows (options 133-135): the new parity test, then iteration counts and
# solve times on the designs real runs reach after 25 iterations (short_cantilever N=512, bridge N=2048).
set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -s -k "cycle_window or cluster_tail or fused_first_two" ) > gpurun_out/r2q_pytest_cycle.txt 2>&1; tail -12 gpurun_out/r2q_pytest_cycle.txt
timeout 600 python tools/cycle_study.py short_cantilever 512 25 "133=4,134=4,135=2" "133=3,134=4,135=2" "133=2,134=4,135=2" "133=1,134=4,135=2" \
   "133=3,134=4,135=3" "119=0" "119=0,133=5,135=2" "119=0,133=4,135=2" "119=0,133=3,135=2" "119=0,133=2,135=2" "119=0,133=5,134=6,135=2" "119=0,133=6,135=2" \
   > gpurun_out/r2q_cycle_study_n512.jsonl 2> gpurun_out/r2q_cycle_study_n512.err; cat gpurun_out/r2q_cycle_study_n512.jsonl | cut -c1-400; tail -3 gpurun_out/r2q_cycle_study_n512.err
timeout 900 python tools/cycle_study.py bridge 2048 25 "133=7,134=7,135=2" "133=6,134=7,135=2" "133=5,134=7,135=2" "133=4,134=7,135=2" "133=3,134=7,135=2" "133=2,134=7,135=2" \
   "133=6,134=7,135=3" "119=0" "119=0,133=8,135=2" "119=0,133=6,135=2" "119=0,133=4,135=2" "119=0,133=8,134=9,135=2" \
   > gpurun_out/r2q_cycle_study_bridge2048.jsonl 2> gpurun_out/r2q_cycle_study_bridge2048.err; cat gpurun_out/r2q_cycle_study_bridge2048.jsonl | cut -c1-400; tail -3 gpurun_out/r2q_cycle_study_bridge2048.err
