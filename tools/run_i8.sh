set -x
timeout 300 python -m pytest tests/test_gpu_nn_i8.py -x -q -m gpu 2>&1 | tail -15
timeout 120 python tools/i8_trace.py -q 2>&1 | tail -20
timeout 120 python tools/i8_trace.py -q --screen 2>&1 | tail -20
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
timeout 200 python bench.py --no-cpu-baseline --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg4 ms/step', d['ms_per_step'], 'nn_ms', d['roofline'].get('nn_kernels_ms'), d['roofline']['frac'], d['selection']); print(d.get('screened'))"
