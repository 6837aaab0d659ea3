"""Diagnostic (GPU box): long random-action rollouts at the BASELINE batch size; every env whose step ended in a
non-finite state (counters[:,4]) is dumped with its PRE-step state (qpos, qvel, warmstart, action, pool row) so that the
step can be replayed in the serial emulation build (tools/build_emu.sh) and in the fp64 oracle.

    python tools/nonfinite_probe.py [--steps 1000] [--envs 4096] [--tasks UnitreeA1.simple ...] [--out gpurun_out/nonfinite]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("LOCO_MUJOCO_B200_FORCE_BUNDLED", "1")
from loco_mujoco_b200 import LocoEnv  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--steps", type=int, default=1000)
    p.add_argument("--envs", type=int, default=4096)
    p.add_argument("--tasks", nargs="*", default=["UnitreeA1.simple", "HumanoidTorque.run", "Atlas.walk", "Talos.walk"])
    p.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "nonfinite"))
    p.add_argument("--seed", type=int, default=0)
    a = p.parse_args()
    os.makedirs(a.out, exist_ok=True)
    summary = {}
    for task in a.tasks:
        N = a.envs
        env = LocoEnv.make(task + ".real", debug=True, num_envs=N, seed=a.seed)
        eng = env._get_engine()
        env.reset()
        g = torch.Generator(device="cuda").manual_seed(a.seed)
        dumps = []
        nd = 0
        prev_bad = eng.counters()[:, 4].clone()
        vmax = 0.0
        for k in range(a.steps):
            act = torch.rand((N, eng.action_dim), device="cuda", generator=g) * 2 - 1
            q, v, w = eng.get_state()
            rows = eng.param_rows()
            obs, rew, done, nxt = eng.step(act)
            nd += int(done.sum())
            c = eng.counters()
            bad = torch.nonzero(c[:, 4] > prev_bad).flatten()
            prev_bad = c[:, 4].clone()
            vmax = max(vmax, float(v.abs().max()))
            for i in bad.tolist():
                dumps.append(dict(step=k, env=i, qpos=q[i].cpu().numpy(), qvel=v[i].cpu().numpy(), ws=w[i].cpu().numpy(),
                                  action=act[i].cpu().numpy(), row=int(rows[i]), iters=int(c[i, 2]), ncon=int(c[i, 3])))
        c = eng.counters().cpu()
        summary[task] = dict(env_steps=int(c[:, 0].sum()), dones=nd, nonfinite=int(c[:, 4].sum()), max_iter_key=int(c[:, 5].max()),
                             max_ncon=int(c[:, 6].max()), max_nefc=int(c[:, 7].max()), max_abs_qvel=vmax)
        print(task, json.dumps(summary[task]), flush=True)
        if dumps:
            np.savez(os.path.join(a.out, task + ".npz"), **{"%s_%d" % (k, j): d[k] for j, d in enumerate(dumps) for k in d})
        del env, eng
    json.dump(summary, open(os.path.join(a.out, "summary.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
