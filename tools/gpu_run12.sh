#!/bin/bash
# round-2 GPU pass 12 (1 GPU): rescore gather A/B (bulk / LSU, ring depth 2 / 3) at C2 and C4 after the full suite
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/r02l_pytest_gpu.log
for cfg in c2 c4; do
  for v in "X=1" "SFB_RESCORE_NST=3" "SFB_RESCORE_LSU=1" "SFB_RESCORE_LSU=1 SFB_RESCORE_NST=3"; do
    echo "== $cfg $v"; env $v timeout 300 python bench.py --config $cfg --no-cpu --no-e2e --no-verify --steps 3 --warmup 2 > gpurun_out/r02l_tmp.json 2>/dev/null; python tools/bench_brief.py gpurun_out/r02l_tmp.json | grep -E "value|knn \{|rescore"
  done
done
