#!/bin/bash
# round-2 GPU pass 5 (1 GPU): C2 with the full-row kNN cross-check, C4, C5 (SF-GRASS + energy lambda + diffusion), launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_cpp_mirror.py -q -m gpu -x 2>&1 | tail -4
timeout 900 python bench.py --verify-full --no-cpu --no-e2e > gpurun_out/r02e_c2_verify_full.json 2> gpurun_out/r02e_c2_verify_full.err
echo "c2 verify-full rc=$?"; python tools/bench_brief.py gpurun_out/r02e_c2_verify_full.json; tail -3 gpurun_out/r02e_c2_verify_full.err
timeout 900 python bench.py --config c4 --steps 5 --warmup 3 > gpurun_out/r02e_c4.json 2> gpurun_out/r02e_c4.err
echo "c4 rc=$?"; python tools/bench_brief.py gpurun_out/r02e_c4.json; tail -3 gpurun_out/r02e_c4.err
timeout 1500 python bench.py --config c5 --steps 2 --warmup 1 --cpu-seconds 10 > gpurun_out/r02e_c5.json 2> gpurun_out/r02e_c5.err
echo "c5 rc=$?"; python tools/bench_brief.py gpurun_out/r02e_c5.json; tail -3 gpurun_out/r02e_c5.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02e_launches_ncu.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-verify > gpurun_out/r02e_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
