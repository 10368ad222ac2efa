set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cluster_tail or state_solve or transfer or mixed" -s 2>&1 | grep -E "iterations without|passed|failed|Error|error" | tail -15
timeout 200 python tools/vcycle_study.py short_cantilever 512 2>&1 | sed -n 4,5p
timeout 300 python bench.py --steps 5 --warmup 3 --no_cpu_baseline > gpurun_out/bench15.json 2> gpurun_out/bench15.err; tail -c 1500 gpurun_out/bench15.err; python -c "
import json;d=json.loads(open('gpurun_out/bench15.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e'],d['gpu_launches'],d['pcg'],d['roofline']['frac'])"
