import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

GOLDEN_TASKS = ["UnitreeA1.simple", "UnitreeA1.hard", "HumanoidTorque.run", "HumanoidTorque.walk", "Atlas.walk", "Talos.walk",
                "UnitreeH1.run", "UnitreeH1.walk", "UnitreeH1.carry", "Atlas.carry", "Talos.carry", "UnitreeG1.run", "UnitreeG1.walk"] + \
    ["HumanoidTorque4Ages.%s.%s" % (t, m) for t in ("run", "walk") for m in "1234"]


# Convex mesh-mesh contacts (HumanoidTorque.walk rows >= 20, HumanoidTorque4Ages run.3 / walk.2-4, UnitreeG1.walk last rows:
# bone against bone, mjc_Convex = libccd MPR) are built since round 2 (oracle: ccd_mpr_penetration; engine:
# mpr_penetration): those goldens are reproduced over the WHOLE episode, same length, to <= 2.6e-6. That residual is the
# MPR tolerance itself (opt.mpr_tolerance = 1e-6: the penetration depth depends at that level on which hull vertices the
# portal visits; scipy's Qhull and MuJoCo's own qhull run do not enumerate identical vertex sets), hence atol 1e-5 below.
# UnitreeH1: (i) the golden's first contact (row 10) is a SELF contact, the right hip-yaw cylinder against the thigh mesh
# (mjc_Convex = MPR, built since round 2; without that pair the row is off by 1.1, with it by 2.8e-3, 1.4e-2 in row 11: the
# contact carries ~800 N and MPR's normal hops between hull facets from sub-step to sub-step, so the last bits of the hull decide);
# (ii) from row 12 on the feet touch down: they are convex MESHES and MuJoCo's plane-mesh routine picks its (up to 3) contact
# vertices by walking the qhull vertex graph of the mesh from the support vertex, an order that cannot be reproduced without
# MuJoCo's own qhull run; the rule the engines use instead (inferred from the UnitreeH1.walk / .carry goldens, see below) cannot
# be checked on this episode any more, because rows 10-11 have already drifted by 1e-2.
# Hence only the 10 rows (90 steps incl. joint limits) before the first contact are pinned.
# UnitreeH1.walk / .carry start in stance: their first rows pin the plane-mesh contact rule (support vertex + the two next-deepest
# vertices at least 0.3 rbound from the first contact; oracle plane_mesh) to 1.5e-6 / 2e-6 (atol 1e-5 below: H1's mesh frames
# carry ~1e-8 of float32 noise); then the swing leg's hip cylinder meets the thigh mesh (row 4 / 3), as in UnitreeH1.run.
PINNED_ROWS = {"UnitreeH1.run": 10, "UnitreeH1.walk": 5, "UnitreeH1.carry": 4}

# Rows of the golden that the FP32 engine is compared on (tests of the CUDA path and of its serial emulation build); the fp64
# oracle is pinned on the whole episodes. MPR's answer is piecewise constant in its inputs (the contact normal is the normal of
# the Minkowski-difference facet the centre ray leaves through): for a shallow, just-touching bone-bone contact an fp32-sized
# difference of the geom poses can select the neighbouring facet (normal a few degrees off), after which an fp32 and an
# fp64 rollout are two different - equally valid - trajectories. In the two longest bone-contact episodes that happens at
# rows 38 / 27 (measured on the emulation build: |obs - golden| jumps from <1e-3 to >1e-2 there); all other goldens are
# followed by the fp32 core over their whole length.
FP32_ROWS = dict(PINNED_ROWS, **{"HumanoidTorque4Ages.walk.2": 36, "HumanoidTorque4Ages.walk.3": 24,
                                 "UnitreeH1.walk": 4, "UnitreeH1.carry": 3})      # (the stance rows: before the hip contact starts)


# Talos.carry: the oracle follows the golden to 2.0e-7 over the whole episode (same episode length / done timing); a few
# near-zero velocity entries miss np.allclose's default atol of 1e-8 (Talos.walk: 5e-8, inside). The residual comes from
# Talos' mesh-derived inertias (float32 STL vertices -> equivalent inertia boxes), not from the dynamics.
GOLDEN_ATOL = {"UnitreeH1.walk": 1e-5, "UnitreeH1.carry": 1e-5, "Talos.carry": 1e-6, "HumanoidTorque.walk": 1e-5, "UnitreeG1.walk": 1e-5, "HumanoidTorque4Ages.run.3": 1e-5,
               "HumanoidTorque4Ages.walk.2": 1e-5, "HumanoidTorque4Ages.walk.3": 1e-5, "HumanoidTorque4Ages.walk.4": 1e-5}


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
                            eps=1e-6, n_probe=4, seed=0, relative=False, user=None):
    """How far the ORACLE's own one-step result moves when its start state is perturbed by `eps` (fp32 resolution).

    A control step is discontinuous where a contact or a joint limit switches on (trajectory samples clipped onto a
    joint limit sit exactly on such a switch) and where MPR changes the facet whose normal it reports for a bone-bone
    contact; there an fp32 engine and an fp64 oracle may legitimately take different branches. Tests use this to tell
    such a state from a real mismatch: the fp32 error must not exceed what an fp32-sized perturbation does to the fp64
    result. relative=True scales the perturbation with the magnitude of each entry (eps * (1 + |x|): random torques drive
    joint velocities to O(100) rad/s, where fp32 resolution is 1e-5, not 1e-7).
    """
    rng = np.random.RandomState(seed)
    gap = 0.0
    for _ in range(n_probe):
        oe = oracle.env(model_blobs, task_blobs)
        if user is not None:
            oe.set_user(user)
        oe.reset_to(int(traj_no), int(step_no))
        sq = (1.0 + np.abs(qpos)) if relative else 1.0
        sv = (1.0 + np.abs(qvel)) if relative else 1.0
        oe.set_state(qpos + eps * sq * rng.randn(len(qpos)), qvel + eps * sv * rng.randn(len(qvel)))
        o, _, _ = oe.step(np.asarray(action, dtype=np.float64))
        gap = max(gap, float(np.abs(o - ref_obs).max()))
        oe.close()
    return gap
