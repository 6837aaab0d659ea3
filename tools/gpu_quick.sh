#!/bin/bash
# quick GPU check (under gpurun): GPU test tier + one short bench per listed task.  usage: bash tools/gpu_quick.sh tag Task1 Task2 ...
TAG=$1; shift
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest.log
for T in "$@"; do
  timeout 300 python bench.py --task $T --steps 60 --warmup 5 --no-configs --no-cpu-baseline 2> $O/${TAG}_$T.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$T', '%.0f env-steps/s, kernel %.3f ms, e2e %.0f, nonfinite %d' % (d['value'], d['kernel_ms_per_step'], d['e2e']['value'], d['nonfinite_in_run']))" | tee -a $O/${TAG}_bench.txt
done
