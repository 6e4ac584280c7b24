#!/bin/bash
# A/B arbitrary build environments on the GPU box: tools/ab_env.sh "VAR=val VAR2=val" "..." ; prints the c2 bench line's key numbers
for E in "$@"; do
  env $E python __graft_entry__.py --force > /dev/null 2>&1 || { echo "build failed: $E"; continue; }
  echo "=== env: '$E'"
  env $E python bench.py --no-cpu-baseline --no-train --no-big 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['stage_ms_per_step'], 'e2e', d['e2e']['value'])"
done
