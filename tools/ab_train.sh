#!/bin/bash
# A/B build environments for the training step on the GPU box: tools/ab_train.sh "VAR=val" ...
for E in "$@"; do
  env $E python __graft_entry__.py --force > /dev/null 2>&1 || { echo "build failed: $E"; continue; }
  echo "=== env: '$E'"
  env $E python tools/diag_train.py 20 2>/dev/null | tail -1 | python -c "
import sys,ast; d=ast.literal_eval(sys.stdin.read()); print(d['ms_per_step'], d['value'])"
done
