#!/bin/bash
# A/B compile-time settings of the trace path on the GPU box: tools/ab_trace.sh "<flags A>" "<flags B>" ...
for F in "$@"; do
  QF_EXTRA_NVCC_FLAGS="$F" python __graft_entry__.py --force > /dev/null 2>&1 || { echo "build failed: $F"; continue; }
  echo "=== flags: '$F'"
  QF_EXTRA_NVCC_FLAGS="$F" python bench.py --no-cpu-baseline --no-train 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['stage_ms_per_step']['trace'], 'dense', d['c2_dense']['ms_per_frame'], 'c4', d['c4']['ms_per_frame'], d['c4']['rank0_stage_ms_per_frame']['trace'], 'c5', d['c5']['ms_per_frame'], d['c5']['rank0_stage_ms_per_frame']['trace'])"
done
