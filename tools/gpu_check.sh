set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "filter" -s 2>&1 | grep -E "^filter|iterations|passed|failed|Error|error|assert" | tail -15
timeout 200 python tools/phase_times.py short_cantilever 512 4 | tail -1
timeout 200 python tools/phase_times.py short_cantilever 512 4 124=4 | tail -1
timeout 200 python tools/phase_times.py short_cantilever 512 4 124=8 | tail -1
timeout 200 python tools/phase_times.py short_cantilever 512 4 124=12 | tail -1
