#!/bin/bash
# round-2 GPU pass 4 (2 GPUs): the NCCL exchange tests, C2 at N=2 with the verify block, forced fine Gram tiles, streamed verify
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multirank.py -q -m gpu -x > gpurun_out/r02d_pytest_2gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02d_pytest_2gpu.log; tail -15 gpurun_out/r02d_pytest_2gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 > gpurun_out/r02d_bench_c2_n2.json 2> gpurun_out/r02d_bench_c2_n2.err
echo "bench c2 n2 rc=$?"; python tools/bench_brief.py gpurun_out/r02d_bench_c2_n2.json; tail -5 gpurun_out/r02d_bench_c2_n2.err
SFB_GRAM_GT=8 timeout 900 $TR bench.py --gpus 2 --steps 2 --warmup 1 --no-e2e > gpurun_out/r02d_bench_c2_n2_gt8.json 2> gpurun_out/r02d_bench_c2_n2_gt8.err
echo "bench c2 n2 gt8 rc=$?"; python tools/bench_brief.py gpurun_out/r02d_bench_c2_n2_gt8.json; tail -5 gpurun_out/r02d_bench_c2_n2_gt8.err
SFB_VERIFY_STREAM=1 timeout 900 $TR bench.py --gpus 2 --config c3 --rows 400000 --steps 1 --warmup 1 > gpurun_out/r02d_bench_c3small_n2.json 2> gpurun_out/r02d_bench_c3small_n2.err
echo "bench c3small n2 rc=$?"; python tools/bench_brief.py gpurun_out/r02d_bench_c3small_n2.json; tail -5 gpurun_out/r02d_bench_c3small_n2.err
timeout 600 python bench.py --no-cpu --no-verify --no-e2e --steps 2 --warmup 2 > gpurun_out/r02d_bench_c2_n1.json 2>/dev/null; python tools/bench_brief.py gpurun_out/r02d_bench_c2_n1.json
