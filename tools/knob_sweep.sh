#!/bin/bash
# diagnostic sweep of the lock-step knobs (under gpurun): prints env-steps/s per setting
O=gpurun_out/sweep_${1:-x}.txt
: > $O
run() { # label, env assignments..., then task
  local label="$1"; shift
  local task="$1"; shift
  local v=$(env "$@" python bench.py --task $task --steps 40 --warmup 5 --no-configs --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.0f %.3f' % (d['value'], d['kernel_ms_per_step']))")
  echo "$task $label: $v" | tee -a $O
}
for T in UnitreeA1.simple HumanoidTorque.run; do
  run "default" $T A=1
  run "GROUP=0 (no iteration lock-step)" $T LOCOSIM_GROUP=0
  run "GROUP=3" $T LOCOSIM_GROUP=3
  run "GROUP=5" $T LOCOSIM_GROUP=5
  run "GROUP=8" $T LOCOSIM_GROUP=8
  run "PHASES=0" $T LOCOSIM_SYNC_PHASES=0
  run "GROUP=0 PHASES=0" $T LOCOSIM_GROUP=0 LOCOSIM_SYNC_PHASES=0
  run "GROUP=0 PHASES=0 SYNC=0" $T LOCOSIM_GROUP=0 LOCOSIM_SYNC_PHASES=0 LOCOSIM_SYNC=0
  run "WPB=8" $T LOCOSIM_WPB=8
  run "WPB=7 (2 blocks/SM)" $T LOCOSIM_WPB=7
  run "WPB=5 (3 blocks/SM)" $T LOCOSIM_WPB=5
  run "WPB=5 GROUP=0" $T LOCOSIM_WPB=5 LOCOSIM_GROUP=0
  run "REGROUP=0" $T LOCOSIM_REGROUP=0
done
