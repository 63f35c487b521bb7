#!/bin/bash
# Round-end style run of one build: GPU test suite, the default bench line, the reference (CPU) arm.  usage: tools/final_run.sh <tag>
tag=${1:-final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${tag}_tests.log
tail -3 gpurun_out/${tag}_tests.log
timeout 900 python bench.py --profile-out gpurun_out/${tag}_profile.json > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/${tag}_bench.json
timeout 600 python bench.py --impl reference > gpurun_out/${tag}_reference.json 2> gpurun_out/${tag}_reference.err; echo "reference rc=$?"
tail -c 400 gpurun_out/${tag}_reference.json
