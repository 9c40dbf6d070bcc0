#!/bin/bash
# round-2 GPU pass 14 (N GPUs, default 2): the multi-rank tests, C2 (default line + trace) and C3
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m pytest tests/test_gpu_multirank.py -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/r02o_pytest_${N}gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
timeout 900 $TR bench.py --gpus $N > gpurun_out/r02o_c2_n$N.json 2> gpurun_out/r02o_c2_n$N.err
echo "c2 n$N rc=$?"; python tools/bench_brief.py gpurun_out/r02o_c2_n$N.json; tail -3 gpurun_out/r02o_c2_n$N.err
SFB_BENCH_TRACE=1 timeout 600 $TR bench.py --gpus $N --steps 2 --warmup 2 --no-e2e --no-verify > /dev/null 2> gpurun_out/r02o_c2_n${N}_trace.err
echo "trace rc=$?"; grep "knn_columns\|allgather" gpurun_out/r02o_c2_n${N}_trace.err | sort | uniq -c | sort -k2,2 -k3n | tail -12
if [ "${2:-1}" = "1" ]; then
  timeout 2400 $TR bench.py --gpus $N --config c3 --steps 2 --warmup 1 --no-e2e > gpurun_out/r02o_c3_n$N.json 2> gpurun_out/r02o_c3_n$N.err
  echo "c3 n$N rc=$?"; python tools/bench_brief.py gpurun_out/r02o_c3_n$N.json; tail -3 gpurun_out/r02o_c3_n$N.err
fi
