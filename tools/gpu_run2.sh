#!/bin/bash
# round-2 GPU pass 2: suite, C2 bench, streamed verify at a reduced C3, ncu of the lambda tile kernel and the Laplacian kernels
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --durations=10 > gpurun_out/r02b_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02b_pytest_gpu.log
tail -40 gpurun_out/r02b_pytest_gpu.log
timeout 900 python bench.py --no-cpu > gpurun_out/r02b_bench_c2.json 2> gpurun_out/r02b_bench_c2.err
echo "bench c2 rc=$?"; python tools/bench_brief.py gpurun_out/r02b_bench_c2.json; tail -5 gpurun_out/r02b_bench_c2.err
SFB_VERIFY_STREAM=1 timeout 900 python bench.py --config c3 --rows 300000 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02b_bench_c3small.json 2> gpurun_out/r02b_bench_c3small.err
echo "bench c3-small rc=$?"; python tools/bench_brief.py gpurun_out/r02b_bench_c3small.json; tail -5 gpurun_out/r02b_bench_c3small.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lambda_tile|lap_merge|csr_emit|rev_scatter|rev_count|scan_lookback' --launch-skip 14 --launch-count 14 \
  -o gpurun_out/r02b_lambda_lap --force-overwrite python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-verify > gpurun_out/r02b_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r02b_ncu.log
