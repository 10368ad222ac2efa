set -x
mkdir -p gpurun_out
for spec in ${NCU_SPECS:-tail:tail_vcycle_kernelId}; do
  name=${spec%%:*}; rx=${spec#*:}
  TM_PROFILER_RANGE=1 timeout 400 ncu --profile-from-start off --set full --import-source on --clock-control none \
    --kernel-name-base mangled -k regex:$rx -c 2 -f -o gpurun_out/${NCU_TAG:-r1h}_$name \
    python bench.py --steps 1 --warmup 3 --no_cpu_baseline --no_e2e --no_mixed_leg > gpurun_out/${NCU_TAG:-r1h}_$name.log 2>&1
  tail -3 gpurun_out/${NCU_TAG:-r1h}_$name.log | cut -c1-300
done
ls -la gpurun_out/*.ncu-rep
