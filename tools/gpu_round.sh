# One gpurun call that validates HEAD: GPU parity tests, the default bench line, the ncu launch
# list of one timed step and `ncu --set full` captures of the top kernels.  Tag = $1.
TAG=${1:-r1h}
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_event_reasons.active --format=csv,noheader
( time timeout 700 python -m pytest tests -x -q -m gpu ) > gpurun_out/${TAG}_pytest_gpu.txt 2>&1; tail -4 gpurun_out/${TAG}_pytest_gpu.txt
timeout 400 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench.err; cut -c1-700 gpurun_out/${TAG}_bench.json
TM_PROFILER_RANGE=1 timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/${TAG}_launches_bench_steps1.csv python bench.py --steps 1 --warmup 3 --no_cpu_baseline --no_e2e --no_mixed_leg > gpurun_out/${TAG}_launches.log 2>&1
wc -l gpurun_out/${TAG}_launches_bench_steps1.csv
for spec in ${NCU_SPECS:-cheb:elast_apply_kernelIdLb0ELi3 filter_tb:filter_cheb_tb_kernel tail:tail_vcycle_kernelId}; do
  name=${spec%%:*}; rx=${spec#*:}
  TM_PROFILER_RANGE=1 timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none \
    --kernel-name-base mangled -k regex:$rx -c 2 -f -o gpurun_out/${TAG}_$name \
    python bench.py --steps 1 --warmup 3 --no_cpu_baseline --no_e2e --no_mixed_leg > gpurun_out/${TAG}_$name.log 2>&1
  tail -2 gpurun_out/${TAG}_$name.log | cut -c1-200
done
ls -la gpurun_out/
