#!/bin/bash
# (under gpurun) A/B of prebuilt engine libraries on ONE box: scratch/variants/*.so are swapped in one after the other
O=gpurun_out/ab_variants.txt; : > $O
cp loco_mujoco_b200/liblocosim_cuda.so /tmp/keep.so
for rep in 1 2; do
for V in scratch/variants/*.so; do
  cp $V loco_mujoco_b200/liblocosim_cuda.so
  for T in ${TASKS:-HumanoidTorque.run Atlas.walk UnitreeA1.simple}; do
    python bench.py --task $T --steps 40 --warmup 5 --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$V $T %.0f %.3f' % (d['value'], d['kernel_ms_per_step']))" | tee -a $O
  done
done
done
cp /tmp/keep.so loco_mujoco_b200/liblocosim_cuda.so
