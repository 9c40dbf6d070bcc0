#!/bin/bash
# round-2 GPU pass 11 (1 GPU): full -m gpu suite; C2 with A/B of the rescore ring depth and the few-row fallback kernel; C4; C5
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/r02k_pytest_gpu.log
echo "== C2"; timeout 600 python bench.py > gpurun_out/r02k_c2.json 2>gpurun_out/r02k_c2.err; python tools/bench_brief.py gpurun_out/r02k_c2.json
for v in SFB_RESCORE_NST2 SFB_EXACT_NO_FEW; do
  echo "== C2 with $v=1"; env $v=1 timeout 300 python bench.py --no-cpu --no-e2e --no-verify --steps 3 --warmup 2 > gpurun_out/r02k_c2_$v.json 2>/dev/null; python tools/bench_brief.py gpurun_out/r02k_c2_$v.json | sed -n 2,4p
done
echo "== C4"; timeout 600 python bench.py --config c4 --no-cpu > gpurun_out/r02k_c4.json 2>gpurun_out/r02k_c4.err; python tools/bench_brief.py gpurun_out/r02k_c4.json
echo "== C5"; timeout 900 python bench.py --config c5 --no-cpu --steps 2 --warmup 1 > gpurun_out/r02k_c5.json 2>gpurun_out/r02k_c5.err; python tools/bench_brief.py gpurun_out/r02k_c5.json
