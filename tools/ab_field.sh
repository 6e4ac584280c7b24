#!/bin/bash
# A/B the field kernel under different compile-time settings on the GPU box: tools/ab_field.sh "<flags A>" "<flags B>" ...
for F in "$@"; do
  QF_EXTRA_NVCC_FLAGS="$F" python __graft_entry__.py --force > /dev/null 2>&1 || { echo "build failed: $F"; continue; }
  echo "=== flags: '$F'"
  QF_EXTRA_NVCC_FLAGS="$F" timeout 200 python tools/diag_variants.py 1 2>&1 | grep variant
  QF_EXTRA_NVCC_FLAGS="$F" python bench.py --no-cpu-baseline --no-train 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['stage_ms_per_step'])"
done
