set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 5 --warmup 3 --no_cpu_baseline > gpurun_out/bench18.json 2> gpurun_out/bench18.err; tail -c 1000 gpurun_out/bench18.err; python -c "
import json;d=json.loads(open('gpurun_out/bench18.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e'],d['gpu_launches'],d['pcg'],d['roofline']['frac'], d['objective_trace'])"
