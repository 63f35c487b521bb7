#!/bin/bash
# N-GPU bench line of the strong-scaled corpus.  usage: tools/final_run_ngpu.sh <tag> <N>
tag=${1:-final}; N=${2:-4}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N \
    > gpurun_out/${tag}_bench_${N}gpu.json 2> gpurun_out/${tag}_bench_${N}gpu.err; echo "bench rc=$?"
tail -c 300 gpurun_out/${tag}_bench_${N}gpu.json
