#!/bin/bash
# round-2 GPU pass 13 (1 GPU): ncu --set full of the rescore kernel at C4 and C2 (one launch each)
mkdir -p gpurun_out
for cfg in c4 c2; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:knn_rescore --launch-skip 1 --launch-count 1 -o gpurun_out/r02m_rescore_$cfg -f \
    python bench.py --config $cfg --steps 1 --warmup 1 --no-cpu --no-e2e --no-verify > /dev/null 2>gpurun_out/r02m_ncu_$cfg.err
  echo "ncu $cfg rc=$?"; ls -la gpurun_out/r02m_rescore_$cfg.ncu-rep
done
