set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
timeout 200 python tools/vcycle_study.py short_cantilever 512 2>&1 | sed -n 1,5p
timeout 300 python bench.py --steps 5 --warmup 3 --no_cpu_baseline > gpurun_out/bench19.json 2> gpurun_out/bench19.err; tail -c 1000 gpurun_out/bench19.err; python -c "
import json;d=json.loads(open('gpurun_out/bench19.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e'],d['gpu_launches'],d['pcg'],d['roofline']['frac']); print({k:(v['ms_sampled']/max(v['launches_sampled'],1)) for k,v in d['roofline']['per_epilogue'].items()})"
