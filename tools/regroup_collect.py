"""Collect (state features at the end of step t) -> (Newton iterations of step t+1) for offline study of regrouping keys. (under gpurun)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("LOCO_MUJOCO_B200_FORCE_BUNDLED", "1")


def main():
    import torch
    from loco_mujoco_b200 import LocoEnv
    task = sys.argv[1]
    N, T, S = 4096, 140, 1024
    env = LocoEnv.make(task + ".real", debug=True, num_envs=N, seed=0, copy_outputs=False)
    eng = env._get_engine()
    eng.reset()
    g = torch.Generator(device=eng.device).manual_seed(0)
    out = dict(it=[], ncon=[], resets=[], qpos=[], qvel=[], act=[])
    for k in range(T):
        a = torch.rand((N, eng.action_dim), device=eng.device, generator=g) * 2 - 1
        eng.step(a, auto_reset=True)
        if k >= 40:
            c = eng.counters()[:S].cpu().numpy()
            q, v, _ = eng.get_state()
            out["it"].append(c[:, 2].copy()); out["ncon"].append(c[:, 3].copy()); out["resets"].append(c[:, 1].copy())
            out["qpos"].append(q[:S].cpu().numpy().astype(np.float32)); out["qvel"].append(v[:S].cpu().numpy().astype(np.float32))
            out["act"].append(a[:S].cpu().numpy().astype(np.float16))
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "regroup_%s.npz" % task), **{k: np.array(v) for k, v in out.items()})
    print(task, "saved", {k: np.array(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
