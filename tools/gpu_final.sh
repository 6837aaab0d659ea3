#!/bin/bash
# final round (under gpurun): GPU test tier, the bench line, launch list, one full capture per headline robot, long-run probes
TAG=${1:-final}
O=gpurun_out; mkdir -p $O
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest.log
timeout 600 python bench.py --steps 100 --warmup 10 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 60 -c 1 -f -o $O/${TAG}_a1 \
    python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-configs > $O/${TAG}_ncu_a1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 60 -c 1 -f -o $O/${TAG}_hum \
    python bench.py --task HumanoidTorque.run --steps 30 --warmup 3 --no-cpu-baseline --no-configs > $O/${TAG}_ncu_hum.log 2>&1
timeout 600 python tools/nonfinite_probe.py --steps 1000 --envs 4096 --out $O/${TAG}_nonfinite > $O/${TAG}_nonfinite.log 2>&1; echo "nonfinite rc=$?"; tail -2 $O/${TAG}_nonfinite.log
timeout 600 python tools/flip_rate.py --envs 256 --steps 60 --out $O/${TAG}_flip_rate.json > $O/${TAG}_flip.log 2>&1; echo "flip rc=$?"; tail -2 $O/${TAG}_flip.log
ls $O | grep ${TAG}
