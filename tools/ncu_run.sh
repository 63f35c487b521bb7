#!/bin/bash
# ncu --set full of the first-scale kernels of each metric (48 pairs 1024x1024, 8 distortions per reference).
#   usage: tools/ncu_run.sh <tag>      -> gpurun_out/<tag>_{butteraugli,ssimulacra2,dssim}.ncu-rep
tag=${1:-prof}
mkdir -p gpurun_out
run() {  # metric, kernel regex, launch count
  cmd="python tools/prof_run.py --metrics $1 --pairs 48 --w 1024 --h 1024 --reps 1 --per-ref 8"
  $cmd > gpurun_out/${tag}_$1_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$2 -c $3 -f -o gpurun_out/${tag}_$1 $cmd > gpurun_out/${tag}_$1_ncu.log 2>&1
  echo "$1 rc=$?"; tail -2 gpurun_out/${tag}_$1_plain.log
}
run butteraugli k_ba_ 12
run ssimulacra2 k_s2_ 5
run dssim k_ds_ 8
