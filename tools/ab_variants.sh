#!/bin/bash
# usage (on the GPU box, from the repo root): tools/ab_variants.sh "bench args" base NAME1 base NAME2 ...
# Runs the bench once per entry with variants/lib_NAME.so swapped in for the product library ("base" = the built one),
# prints step / screen / rescore ms, and restores the product library.  Development aid.
args=$1; shift
L=matternet-rs_b200/libsurfface_b200.so
cp $L /tmp/sfb_base.so
for v in "$@"; do
  if [ "$v" = base ]; then cp /tmp/sfb_base.so $L; else cp variants/lib_$v.so $L; fi
  timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e $args 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['knn']
print('$v', 'step', round(d['ms_per_step'],1), 'screen', round(k['ms_screen'],1), 'rescore', round(k['ms_rescore'],1), 'certified', k['rows_certified'], 'fallback', k['rows_fallback'], 'sm_mhz', d['clocks']['sm_mhz'])"
done
cp /tmp/sfb_base.so $L
