#!/bin/bash
# usage: tools/sweep_env.sh VAR "bench args" v1 v2 ...   -- runs the bench (no e2e / cpu) once per value of the env var
var=$1; shift
args=$1; shift
for v in "$@"; do
  env $var=$v timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e $args 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['knn']
print('$var=$v', 'kprime', k['k_prime'], 'step', round(d['ms_per_step'],1), 'screen', round(k['ms_screen'],1), 'rescore', round(k['ms_rescore'],1), 'rescreened', k['rows_rescreened'], round(k['ms_rescreen'],1), 'fallback', k['rows_fallback'], round(k['ms_fallback'],1))"
done
