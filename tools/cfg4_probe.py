"""Probe of BASELINE config 4 (Atlas.walk + Talos.walk with domain randomisation, 1024 + 1024 envs on one GPU): launch
geometry variants (warps per block per engine) -> env-steps/s.   (under gpurun)   python tools/cfg4_probe.py"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("LOCO_MUJOCO_B200_FORCE_BUNDLED", "1")
import bench  # noqa: E402


def main():
    import torch
    a = argparse.Namespace(steps=40, warmup=5, no_flush=False, gather_chunk=0)
    dr = lambda robot: "domain_randomization_%s.yaml" % robot
    members = [("Atlas.walk", 1024, {"domain_randomization_config": dr("atlas")}),
               ("Talos.walk", 1024, {"domain_randomization_config": dr("talos")})]
    for label, wpbs in [("default", (None, None)), ("7/7", (7, 7)), ("7/15", (7, None)), ("5/5", (5, 5)), ("7/8", (7, 8)), ("4/5", (4, 5)),
                        ("10/5", (10, 5)), ("14/7", (None, 7))]:
        mem = [(t, n, dict(kw, warps_per_block=w)) for (t, n, kw), w in zip(members, wpbs)]
        wl = bench.Workload("cfg4", mem, 0, 1, 0)
        r = wl.measure(a, 0, False)
        print(label, [e.launch_info() for e in wl.engines], "%.0f env-steps/s, %.3f ms" % (r["value"], r["ms_per_step"]), flush=True)
        del wl
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
