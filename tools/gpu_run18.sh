#!/bin/bash
# round-2: the C5 line with three warm-up steps and the final kernels
mkdir -p gpurun_out
timeout 600 python bench.py --config c5 --no-cpu --steps 2 --warmup 3 > gpurun_out/r02r_c5.json 2>gpurun_out/r02r_c5.err
echo "rc=$?"; python tools/bench_brief.py gpurun_out/r02r_c5.json | cut -c1-400 | head -16
