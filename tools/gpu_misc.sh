#!/bin/bash
# (under gpurun) racecheck / memcheck of small steps + an A/B of the regrouping defaults
O=gpurun_out; mkdir -p $O
cat > /tmp/san_step.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
os.environ.setdefault("LOCO_MUJOCO_B200_FORCE_BUNDLED", "1")
from loco_mujoco_b200 import LocoEnv
task = sys.argv[1]
env = LocoEnv.make(task + ".real", debug=True, num_envs=64, seed=3)
obs = env.reset()
g = torch.Generator(device="cpu").manual_seed(0)
for k in range(int(sys.argv[2])):
    a = (torch.rand((64, env.info.action_space.shape[0]), generator=g) * 2 - 1).cuda()
    obs, r, d, _ = env.step(a)
torch.cuda.synchronize()
print(task, "ok", float(obs.abs().max()))
PY
for T in UnitreeA1.simple HumanoidTorque.run; do
  timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python /tmp/san_step.py $T 3 > $O/racecheck_$T.log 2>&1; echo "racecheck $T rc=$?"; tail -3 $O/racecheck_$T.log
  timeout 600 compute-sanitizer --tool memcheck python /tmp/san_step.py $T 2 > $O/memcheck_$T.log 2>&1; echo "memcheck $T rc=$?"; tail -3 $O/memcheck_$T.log
done
run() { local label="$1"; shift; local task="$1"; shift
  local v=$(env "$@" python bench.py --task $task --steps 60 --warmup 5 --no-configs --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.0f %.3f' % (d['value'], d['kernel_ms_per_step']))")
  echo "$task $label: $v" | tee -a $O/sweep_c.txt
}
: > $O/sweep_c.txt
for rep in 1 2; do
  run "W=32 KR=0 (default)" UnitreeA1.simple A=1
  run "W=32 KR=off" UnitreeA1.simple LOCOSIM_KEY_RESET=-1
  run "W=4 KR=0" UnitreeA1.simple LOCOSIM_MPR_WEIGHT=4
  run "W=4 KR=off" UnitreeA1.simple LOCOSIM_MPR_WEIGHT=4 LOCOSIM_KEY_RESET=-1
  run "W=0 KR=off" UnitreeA1.simple LOCOSIM_MPR_WEIGHT=0 LOCOSIM_KEY_RESET=-1
done
run "default" UnitreeG1.run A=1
run "default" UnitreeH1.run A=1
run "default" Talos.walk A=1
run "KR=0" Talos.walk LOCOSIM_KEY_RESET=0
run "default" Atlas.walk A=1
run "KR=0" Atlas.walk LOCOSIM_KEY_RESET=0
