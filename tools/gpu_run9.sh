#!/bin/bash
# round-2 GPU pass 9 (1 GPU): Gram side-job modes at N = 1, rescore counters
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "columns or pipeline or c1" 2>&1 | tail -3
for mode in co after; do
  echo "== SFB_GRAM_MODE=$mode"; SFB_GRAM_MODE=$mode timeout 300 python bench.py --no-cpu --no-e2e --no-verify --steps 3 --warmup 2 > gpurun_out/r02i_c2_$mode.json 2>/dev/null; python tools/bench_brief.py gpurun_out/r02i_c2_$mode.json | head -4
  SFB_BENCH_TRACE=1 SFB_GRAM_MODE=$mode timeout 300 python bench.py --no-cpu --no-e2e --no-verify --steps 1 --warmup 1 2>&1 >/dev/null | grep "knn \|knn_columns\|adjacency \|laplacian \|lambda " | tail -5
done
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,smsp__inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:knn_rescore --launch-skip 1 --launch-count 1 --csv --log-file gpurun_out/r02i_rescore_metrics.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-verify > /dev/null 2>&1
echo "ncu rescore rc=$?"; tail -8 gpurun_out/r02i_rescore_metrics.csv | cut -c1-300
