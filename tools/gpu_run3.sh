#!/bin/bash
# round-2 GPU pass 3: suite, C2 bench, eight-warp epilogue probes, ncu of lambda + Laplacian kernels
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --durations=5 > gpurun_out/r02c_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02c_pytest_gpu.log
tail -25 gpurun_out/r02c_pytest_gpu.log
timeout 900 python bench.py --no-cpu > gpurun_out/r02c_bench_c2.json 2> gpurun_out/r02c_bench_c2.err
echo "bench c2 rc=$?"; python tools/bench_brief.py gpurun_out/r02c_bench_c2.json; tail -5 gpurun_out/r02c_bench_c2.err
for dbg in 2 1 3; do
  for ew in 0 1; do
    echo "== EW8=$ew DBG=$dbg"; SFB_SCREEN_EW8=$ew SFB_SCREEN_DBG=$dbg timeout 120 python bench.py --no-cpu --no-e2e --no-verify --steps 2 --warmup 1 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('   screen ms', [s[1] for s in d['knn_ms_steps']], 'step', round(d['ms_per_step'],1))
"
  done
done
echo "== EW8=1 full"; SFB_SCREEN_EW8=1 timeout 300 python bench.py --no-cpu --no-e2e --steps 3 --warmup 2 > gpurun_out/r02c_bench_ew8.json 2> gpurun_out/r02c_bench_ew8.err; python tools/bench_brief.py gpurun_out/r02c_bench_ew8.json; tail -3 gpurun_out/r02c_bench_ew8.err
SFB_SCREEN_EW8=1 timeout 600 python -m pytest tests/test_gpu_screen.py -q -m gpu -x 2>&1 | tail -5
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lambda_tile|lap_merge_rows|csr_copy|rev_scatter' --launch-skip 4 --launch-count 4 \
  -o gpurun_out/r02c_lambda_lap --force-overwrite python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-verify > gpurun_out/r02c_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r02c_ncu.log | cut -c1-300
