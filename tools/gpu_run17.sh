#!/bin/bash
# round-2: every screened row of C2 against the exact f64 kernel, with the final rescore / fallback kernels
mkdir -p gpurun_out
timeout 600 python bench.py --verify-full --no-e2e --no-cpu > gpurun_out/r02q_c2_verify_full.json 2>gpurun_out/r02q_c2_verify_full.err
echo "rc=$?"; python tools/bench_brief.py gpurun_out/r02q_c2_verify_full.json | grep -E "value|verify" | cut -c1-900
