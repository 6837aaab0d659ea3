#!/bin/bash
# regrouping-key sweep for the convex (MPR) workload (under gpurun)
O=gpurun_out/sweep_${1:-y}.txt
: > $O
run() { local label="$1"; shift; local task="$1"; shift
  local v=$(env "$@" python bench.py --task $task --steps 40 --warmup 5 --no-configs --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.0f %.3f' % (d['value'], d['kernel_ms_per_step']))")
  echo "$task $label: $v" | tee -a $O
}
T=HumanoidTorque.run
run "default" $T A=1
for W in 8 16 32; do run "MPR_WEIGHT=$W" $T LOCOSIM_MPR_WEIGHT=$W; done
run "KEY_RESET=0" $T LOCOSIM_KEY_RESET=0
run "KEY_RESET=0 W=16" $T LOCOSIM_KEY_RESET=0 LOCOSIM_MPR_WEIGHT=16
run "KEY_RESET=40 W=16" $T LOCOSIM_KEY_RESET=40 LOCOSIM_MPR_WEIGHT=16
run "KEY_SHIFT=3 W=16" $T LOCOSIM_KEY_SHIFT=3 LOCOSIM_MPR_WEIGHT=16
run "KEY_SHIFT=3 W=32 RESET=0" $T LOCOSIM_KEY_SHIFT=3 LOCOSIM_MPR_WEIGHT=32 LOCOSIM_KEY_RESET=0
run "WPB=7" $T LOCOSIM_WPB=7
run "WPB=7 W=16" $T LOCOSIM_WPB=7 LOCOSIM_MPR_WEIGHT=16
run "no convex (debug 4)" $T LOCOSIM_DEBUG=4
run "no MPR (debug 1)" $T LOCOSIM_DEBUG=1
run "no OBB filter+no MPR (debug 3)" $T LOCOSIM_DEBUG=3
run "A1 default" UnitreeA1.simple A=1
run "A1 KEY_RESET=0" UnitreeA1.simple LOCOSIM_KEY_RESET=0
run "A1 no convex (debug 4)" UnitreeA1.simple LOCOSIM_DEBUG=4
