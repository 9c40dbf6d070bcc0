#!/bin/bash
# round-2 GPU pass 10 (1 GPU): full -m gpu suite after the register-ring Gram kernel and the narrow lambda tiles; Gram probe; C2 + C4 lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/r02j_pytest_gpu.log
echo "== gram probe"; timeout 900 python tools/gram_probe.py 2>&1 | tee gpurun_out/r02j_gram_probe.txt
echo "== C2"; timeout 600 python bench.py > gpurun_out/r02j_c2.json 2>gpurun_out/r02j_c2.err; python tools/bench_brief.py gpurun_out/r02j_c2.json
SFB_BENCH_TRACE=1 timeout 300 python bench.py --no-cpu --no-e2e --no-verify --steps 1 --warmup 1 2>&1 >/dev/null | grep "knn \|knn_columns\|adjacency \|laplacian \|lambda " | tail -5
echo "== C4"; timeout 600 python bench.py --config c4 --no-cpu > gpurun_out/r02j_c4.json 2>gpurun_out/r02j_c4.err; python tools/bench_brief.py gpurun_out/r02j_c4.json
