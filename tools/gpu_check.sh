set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cluster_tail" -s 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python bench.py --steps 5 --warmup 3 --no_cpu_baseline > gpurun_out/bench13.json 2> gpurun_out/bench13.err; tail -c 1500 gpurun_out/bench13.err; python -c "
import json;d=json.loads(open('gpurun_out/bench13.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e'],d['gpu_launches'],d['pcg'])"
timeout 300 python bench.py --steps 5 --warmup 3 --no_cpu_baseline --no_e2e --engine_option 119=0 > gpurun_out/bench13_notail.json 2>> gpurun_out/bench13.err; python -c "
import json;d=json.loads(open('gpurun_out/bench13_notail.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['gpu_launches'],d['pcg'])"
timeout 300 python bench.py --steps 5 --warmup 3 --no_cpu_baseline --no_e2e --engine_option 120=8 > gpurun_out/bench13_c8.json 2>> gpurun_out/bench13.err; python -c "
import json;d=json.loads(open('gpurun_out/bench13_c8.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['gpu_launches'],d['pcg'])"
timeout 300 python bench.py --steps 5 --warmup 3 --no_cpu_baseline --no_e2e --engine_option 119=9000 > gpurun_out/bench13_t9000.json 2>> gpurun_out/bench13.err; python -c "
import json;d=json.loads(open('gpurun_out/bench13_t9000.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['gpu_launches'],d['pcg'])"
