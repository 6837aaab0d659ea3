import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

GOLDEN_TASKS = ["UnitreeA1.simple", "UnitreeA1.hard", "HumanoidTorque.run", "HumanoidTorque.walk", "Atlas.walk", "Talos.walk"]


# HumanoidTorque.walk: from row 20 on the reference rollout contains a convex mesh-mesh self-contact (fixed arm/hand
# bones against the leg, mjc_Convex / libccd MPR in MuJoCo) that the engines do not implement yet (DESIGN.md "gaps");
# rows 0..19 (190 RK4 steps = 760 dynamics evaluations) are pinned.
PINNED_ROWS = {"HumanoidTorque.walk": 20}


def golden(task):
    return np.load(os.path.join(GOLDEN, task + ".real.npy"))


def make_env(task, **kw):
    from loco_mujoco_b200 import LocoEnv
    return LocoEnv.make(task + ".real", debug=True, **kw)


def blobs(env):
    from loco_mujoco_b200 import modelpack
    return modelpack.pack(env._model), env.task_spec().pack()


def reference_draws(env, seed=0):
    """Replay the legacy numpy RNG stream of the reference test (tests/test_environments.py:15-38,76):
    seed -> reset draws (model idx, traj, sample) -> one randn(nu)*0.1 per step."""
    np.random.seed(seed)
    np.random.randint(0, 1)
    traj_no = np.random.randint(0, env.trajectories.number_of_trajectories)
    step_no = np.random.randint(0, env.trajectories.trajectory_length)
    return traj_no, step_no
