"""How much of the lock-step waiting could a better regrouping key remove?  (under gpurun)
Per control step the engine reports every env's Newton iterations (counters[:,2]) and its key (counters[:,5], computed from
the PREVIOUS step).  Block cost ~ max over the block's envs of the iterations; compare the grouping the engine used (sorted by
its key) with a random grouping and with the unattainable perfect one (sorted by the step's own iteration count)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("LOCO_MUJOCO_B200_FORCE_BUNDLED", "1")


def main():
    import torch
    from loco_mujoco_b200 import LocoEnv
    task = sys.argv[1] if len(sys.argv) > 1 else "UnitreeA1.simple"
    N, T = 4096, 160
    env = LocoEnv.make(task + ".real", debug=True, num_envs=N, seed=0, copy_outputs=False)
    eng = env._get_engine()
    W = eng.launch_info()["warps_per_block"]
    eng.reset()
    g = torch.Generator(device=eng.device).manual_seed(0)
    key_prev = None
    rows = []
    for k in range(T):
        a = torch.rand((N, eng.action_dim), device=eng.device, generator=g) * 2 - 1
        eng.step(a, auto_reset=True)
        c = eng.counters().cpu().numpy()
        it, key = c[:, 2].astype(np.int64), c[:, 5].astype(np.int64)
        if key_prev is not None and k >= 60:
            def cost(order):
                o = it[order]
                pad = (-len(o)) % W
                o = np.concatenate([o, np.zeros(pad, dtype=o.dtype)]).reshape(-1, W)
                return o.max(axis=1).sum() * W
            used = np.argsort(-key_prev, kind="stable")
            rnd = np.random.RandomState(k).permutation(N)
            perfect = np.argsort(-it, kind="stable")
            rows.append((it.sum(), cost(used), cost(rnd), cost(perfect), np.corrcoef(key_prev, it)[0, 1]))
        key_prev = key
    r = np.array(rows, dtype=np.float64)
    print(task, "warps per block", W, "steps", len(r))
    print("mean Newton iterations per env-step: %.1f" % (r[:, 0].mean() / N))
    print("lock-step cost (sum over blocks of W x max iterations) relative to the useful iterations:")
    print("  random grouping %.2f   engine's key (previous step) %.2f   perfect foresight %.2f   corr(key, iterations) %.2f"
          % ((r[:, 2] / r[:, 0]).mean(), (r[:, 1] / r[:, 0]).mean(), (r[:, 3] / r[:, 0]).mean(), r[:, 4].mean()))


if __name__ == "__main__":
    main()
