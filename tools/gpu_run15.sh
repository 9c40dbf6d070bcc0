#!/bin/bash
# round-2 GPU pass 15 (1 GPU): suite; rescore chunk A/B at C2; C4 with 32-edge Gram tiles; C5 with the 8-warp screen epilogue
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/r02n_pytest_gpu.log
for v in "X=1" "SFB_RESCORE_CH=16"; do
  echo "== c2 $v"; env $v timeout 300 python bench.py --no-cpu --no-e2e --no-verify --steps 3 --warmup 2 > gpurun_out/r02n_tmp.json 2>/dev/null; python tools/bench_brief.py gpurun_out/r02n_tmp.json | grep -E "value|knn \{|rescore"
done
for v in "X=1" "SFB_GRAM_GT=16" "SFB_RESCORE_CH=16"; do
  echo "== c4 $v"; env $v timeout 300 python bench.py --config c4 --no-cpu --no-e2e --steps 3 --warmup 2 > gpurun_out/r02n_c4_$v.json 2>/dev/null; python tools/bench_brief.py gpurun_out/r02n_c4_$v.json | grep -E "value|knn \{|rescore|verify"
done
SFB_BENCH_TRACE=1 timeout 300 python bench.py --config c4 --no-cpu --no-e2e --no-verify --steps 1 --warmup 1 2>&1 >/dev/null | grep "knn \|knn_columns\|adjacency \|laplacian \|lambda " | tail -5
for v in "SFB_SCREEN_EW8=1" "SFB_RESCORE_CH=16"; do
  echo "== c5 $v"; env $v timeout 600 python bench.py --config c5 --no-cpu --no-e2e --steps 1 --warmup 1 > gpurun_out/r02n_c5_$v.json 2>/dev/null; python tools/bench_brief.py gpurun_out/r02n_c5_$v.json | grep -E "value|knn \{|roofline|rescore|verify"
done
