# Round 2, call S (1 GPU): automatic cycle window (option 135 = 0) against the V-cycle, coarse smoothing degree 2 and
# a larger exactly solved coarsest level under the W window.
set -x
mkdir -p gpurun_out
timeout 600 python tools/cycle_study.py short_cantilever 512 25 "135=1" "109=2" "4=4" "109=2,4=4" "109=2,133=4,134=6,135=2" "109=2,133=5,134=7,135=2" "109=2,133=4,134=7,135=2" "109=2,133=5,134=6,135=3" "3=4" "3=4,109=2" \
   > gpurun_out/r2s_cycle_study_n512.jsonl 2> gpurun_out/r2s_cycle_study_n512.err; cat gpurun_out/r2s_cycle_study_n512.jsonl | cut -c1-300; tail -3 gpurun_out/r2s_cycle_study_n512.err
timeout 900 python tools/cycle_study.py bridge 2048 25 "135=1" "109=2" "4=4" "109=2,4=4" "109=2,133=5,134=9,135=2" "109=2,133=4,134=9,135=2" "109=2,133=6,134=9,135=3" "133=6,134=8,135=2" "133=7,134=8,135=2" \
   > gpurun_out/r2s_cycle_study_bridge2048.jsonl 2> gpurun_out/r2s_cycle_study_bridge2048.err; cat gpurun_out/r2s_cycle_study_bridge2048.jsonl | cut -c1-300; tail -3 gpurun_out/r2s_cycle_study_bridge2048.err
