#!/bin/bash
# A/B run on the GPU box: the GPU test suite on the tree's library, then a short corpus bench (per-kernel table) of the
# tree's library and of every experiment build under codec_eval_b200/build/exp_*.so.   usage: tools/ab_run.sh <tag>
tag=${1:-ab}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/${tag}_tests.log
tail -5 gpurun_out/${tag}_tests.log
run() {   # name, CE_LIB_PATH or ""
  CE_LIB_PATH=$2 timeout 600 python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline --no-e2e \
      --profile-out gpurun_out/${tag}_$1_profile.json > gpurun_out/${tag}_$1_bench.json 2> gpurun_out/${tag}_$1.err
  echo "$1 rc=$?"; python tools/show_bench.py gpurun_out/${tag}_$1_bench.json 2>/dev/null | head -40
}
run main ""
for so in codec_eval_b200/build/exp_*.so; do
  [ -e "$so" ] || continue
  n=$(basename $so .so); run ${n#exp_} $PWD/$so
done
