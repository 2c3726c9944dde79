#!/bin/bash
# parity tests + stage times of the bench (no CPU baseline)
python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py -m gpu -q -x 2>&1 | tail -3
python bench.py --steps 50 --warmup 5 --no-cpu-baseline "$@" | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value', round(d['value']), 'utt/s  ms/step', round(d['ms_per_step'], 4), ' e2e', round(d['e2e']['value']), ' clocks', d['clocks']['sm_mhz'])
for k, v in d['stages'].items():
    print(f'  {k:28s} {v[\"ms_per_step\"]:.4f} ms', v.get('tflops', ''))
print('frontend', d['frontend_roofline']['achieved'], 'GB/s')
"
