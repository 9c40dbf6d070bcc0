#!/bin/bash
# usage: tools/sweep_env.sh VAR v1 v2 ...   -- runs the C2 bench (no e2e / cpu) once per value of the env var
var=$1; shift
for v in "$@"; do
  env $var=$v timeout 200 python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['knn']
print('$var=$v', 'kprime', k['k_prime'], 'step', round(d['ms_per_step'],1), 'screen', round(k['ms_screen'],1), 'rescore', round(k['ms_rescore'],1), 'rescreened', k['rows_rescreened'], 'fallback', k['rows_fallback'])"
done
