#!/bin/bash
# timing-only experiments on the GRU recurrence: pieces of the per-step chain switched off (results are garbage)
for d in 0 1 2 4 8 16 9 13 29 31; do
  SIR_GRU_DEBUG_SKIP=$d python bench.py --steps 30 --warmup 5 --no-cpu-baseline | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('dbg', $d, ' gru_l0', d['stages']['gru_l0_recurrence']['ms_per_step'], ' gru_l1', d['stages']['gru_l1_recurrence']['ms_per_step'])"
done
