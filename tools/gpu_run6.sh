#!/bin/bash
# round-2 GPU pass 6 (8 GPUs): C2 at N=8 (default line + call trace + Gram tile A/B), C3 = 10M x 768 L2 at N=8 with the streamed verify
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
timeout 900 $TR bench.py --gpus $N > gpurun_out/r02f_c2_n$N.json 2> gpurun_out/r02f_c2_n$N.err
echo "c2 n$N rc=$?"; python tools/bench_brief.py gpurun_out/r02f_c2_n$N.json; tail -3 gpurun_out/r02f_c2_n$N.err
SFB_BENCH_TRACE=1 timeout 600 $TR bench.py --gpus $N --steps 2 --warmup 2 --no-e2e --no-verify > gpurun_out/r02f_c2_n${N}_trace.json 2> gpurun_out/r02f_c2_n${N}_trace.err
echo "trace rc=$?"; grep -A14 "^  knn " gpurun_out/r02f_c2_n${N}_trace.err | tail -40
SFB_GRAM_GT=16 timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 2 --no-e2e --no-verify > gpurun_out/r02f_c2_n${N}_gt16.json 2> gpurun_out/r02f_c2_n${N}_gt16.err
echo "gt16 rc=$?"; python tools/bench_brief.py gpurun_out/r02f_c2_n${N}_gt16.json | head -12
timeout 1800 $TR bench.py --gpus $N --config c3 --steps 3 --warmup 1 --no-e2e > gpurun_out/r02f_c3_n$N.json 2> gpurun_out/r02f_c3_n$N.err
echo "c3 n$N rc=$?"; python tools/bench_brief.py gpurun_out/r02f_c3_n$N.json; tail -3 gpurun_out/r02f_c3_n$N.err
