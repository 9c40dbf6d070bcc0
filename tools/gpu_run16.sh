#!/bin/bash
# round-2 final pass (1 GPU): suite, smoke, the bench lines of C2 (default) / C1 / C4 [/ C5], launch list of the default command
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r02p_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== C2"; timeout 600 python bench.py > gpurun_out/r02p_c2.json 2>gpurun_out/r02p_c2.err; python tools/bench_brief.py gpurun_out/r02p_c2.json
echo "== C1"; timeout 300 python bench.py --config c1 > gpurun_out/r02p_c1.json 2>gpurun_out/r02p_c1.err; python tools/bench_brief.py gpurun_out/r02p_c1.json | head -8
echo "== C4"; timeout 600 python bench.py --config c4 > gpurun_out/r02p_c4.json 2>gpurun_out/r02p_c4.err; python tools/bench_brief.py gpurun_out/r02p_c4.json | head -12
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02p_launches_ncu.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-verify > /dev/null 2>&1
echo "ncu launch list rc=$?"; wc -l gpurun_out/r02p_launches_ncu.csv
if [ "${1:-0}" = "1" ]; then
  echo "== C5"; timeout 900 python bench.py --config c5 --no-cpu --steps 2 --warmup 3 > gpurun_out/r02p_c5.json 2>gpurun_out/r02p_c5.err; python tools/bench_brief.py gpurun_out/r02p_c5.json | head -14
fi
