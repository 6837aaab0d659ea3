#!/bin/bash
# (under gpurun) one full ncu capture per listed task.  usage: bash tools/gpu_ncu.sh tag Task1 [Task2 ...]
TAG=$1; shift
O=gpurun_out; mkdir -p $O
for T in "$@"; do
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 60 -c 1 -f -o $O/${TAG}_$T \
    python bench.py --task $T --steps 30 --warmup 3 --no-cpu-baseline --no-configs > $O/${TAG}_ncu_$T.log 2>&1
  echo "$T rc=$?"
done
ls -la $O | grep ncu-rep
