#!/bin/bash
# A/B compile-time settings of the training kernels on the GPU box: tools/ab_bwd.sh "<flags A>" "<flags B>" ...
for F in "$@"; do
  QF_EXTRA_NVCC_FLAGS="$F" python __graft_entry__.py --force > /dev/null 2>&1 || { echo "build failed: $F"; continue; }
  echo "=== flags: '$F'"
  QF_EXTRA_NVCC_FLAGS="$F" python tools/diag_bwd_time.py 2>&1 | grep bwd_time
  QF_EXTRA_NVCC_FLAGS="$F" python tools/diag_train.py 20 2>/dev/null | tail -1 | python -c "
import sys,ast; d=ast.literal_eval(sys.stdin.read()); print('train step', d['ms_per_step'], d['value'])"
done
