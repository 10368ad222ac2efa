set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python bench.py --steps 5 --warmup 3 --no_cpu_baseline > gpurun_out/bench12.json 2> gpurun_out/bench12.err; tail -c 3000 gpurun_out/bench12.json
timeout 300 python bench.py --steps 5 --warmup 3 --no_cpu_baseline --engine_option 118=0 > gpurun_out/bench12_nofuse.json 2>> gpurun_out/bench12.err; tail -c 600 gpurun_out/bench12_nofuse.json
