#!/bin/bash
# 2-GPU evidence: the NCCL shard test (skips itself below 2 GPUs) and the 2-rank bench line.  usage: tools/final_run_2gpu.sh <tag>
tag=${1:-final}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_shard_gpu.py -m gpu -q > gpurun_out/${tag}_nccl_2gpu_test.log 2>&1; echo "shard test rc=$?"; tail -2 gpurun_out/${tag}_nccl_2gpu_test.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 \
    > gpurun_out/${tag}_bench_2gpu.json 2> gpurun_out/${tag}_bench_2gpu.err; echo "bench rc=$?"
tail -c 300 gpurun_out/${tag}_bench_2gpu.json
