#!/bin/bash
# round-2 GPU pass 1: the -m gpu suite, the default bench (C2, verify on) and C1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_gpus.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu -x --durations=15 > gpurun_out/r02a_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02a_pytest_gpu.log
tail -30 gpurun_out/r02a_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02a_bench_c2.json 2> gpurun_out/r02a_bench_c2.err
echo "bench c2 rc=$?"; tail -c 3000 gpurun_out/r02a_bench_c2.json; tail -5 gpurun_out/r02a_bench_c2.err
timeout 600 python bench.py --config c1 --steps 5 --warmup 3 > gpurun_out/r02a_bench_c1.json 2> gpurun_out/r02a_bench_c1.err
echo "bench c1 rc=$?"; tail -c 1500 gpurun_out/r02a_bench_c1.json; tail -5 gpurun_out/r02a_bench_c1.err
