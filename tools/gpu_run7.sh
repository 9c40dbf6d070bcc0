#!/bin/bash
# round-2 GPU pass 7 (1 GPU): suite, accumulator stress ratios, C2, C5 after the Laplacian / lambda / diffusion changes, full ncu capture of the screen kernel
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --durations=5 > gpurun_out/r02g_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02g_pytest_gpu.log; tail -12 gpurun_out/r02g_pytest_gpu.log
timeout 600 python -m pytest tests/test_gpu_screen.py -q -m gpu -s -k "adversarial and tile" 2>&1 | grep "ACCUM" > gpurun_out/r02g_accum.txt; sort -t= -k5 -n gpurun_out/r02g_accum.txt | tail -6
timeout 900 python bench.py --no-cpu > gpurun_out/r02g_c2.json 2> gpurun_out/r02g_c2.err
echo "c2 rc=$?"; python tools/bench_brief.py gpurun_out/r02g_c2.json; tail -3 gpurun_out/r02g_c2.err
timeout 1500 python bench.py --config c5 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02g_c5.json 2> gpurun_out/r02g_c5.err
echo "c5 rc=$?"; python tools/bench_brief.py gpurun_out/r02g_c5.json; tail -3 gpurun_out/r02g_c5.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'knn_screen_pair|knn_rescore' --launch-skip 2 --launch-count 2 \
  -o gpurun_out/r02g_screen --force-overwrite python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-verify > gpurun_out/r02g_ncu_screen.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r02g_ncu_screen.log | cut -c1-200
