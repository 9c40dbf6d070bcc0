#!/bin/bash
# round-2 GPU pass 8 (N GPUs): C2 default line (+ trace), optionally C3
mkdir -p gpurun_out
N=${1:-8}; WITH_C3=${2:-0}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
timeout 900 $TR bench.py --gpus $N > gpurun_out/r02h_c2_n$N.json 2> gpurun_out/r02h_c2_n$N.err
echo "c2 n$N rc=$?"; python tools/bench_brief.py gpurun_out/r02h_c2_n$N.json > gpurun_out/r02h_c2_n$N.txt; cat gpurun_out/r02h_c2_n$N.txt; tail -3 gpurun_out/r02h_c2_n$N.err
SFB_BENCH_TRACE=1 timeout 600 $TR bench.py --gpus $N --steps 2 --warmup 2 --no-e2e --no-verify > gpurun_out/r02h_c2_n${N}_trace.json 2> gpurun_out/r02h_c2_n${N}_trace.err
echo "trace rc=$?"; grep "knn_columns\|allgather\|^  knn  \|laplacian  \|lambda  " gpurun_out/r02h_c2_n${N}_trace.err | sort | uniq -c | sort -k2,2 -k3n | awk '{print}' | tail -40
if [ "$WITH_C3" = "1" ]; then
  timeout 2400 $TR bench.py --gpus $N --config c3 --steps 3 --warmup 1 --no-e2e > gpurun_out/r02h_c3_n$N.json 2> gpurun_out/r02h_c3_n$N.err
  echo "c3 n$N rc=$?"; python tools/bench_brief.py gpurun_out/r02h_c3_n$N.json > gpurun_out/r02h_c3_n$N.txt; cat gpurun_out/r02h_c3_n$N.txt; tail -3 gpurun_out/r02h_c3_n$N.err
fi
