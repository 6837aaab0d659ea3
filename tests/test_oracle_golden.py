"""
Pins the CPU oracle (oracle/locosim_ref.c) against the reference's own golden rollouts
(/root/reference/tests/test_datasets/*.npy, copied verbatim to tests/golden/): seeded reset + randn*0.1 actions
until has_fallen, compared with np.allclose exactly like /root/reference/tests/test_environments.py:88-94.
"""
import numpy as np
import pytest

from helpers import GOLDEN_TASKS, PINNED_ROWS, GOLDEN_ATOL, golden, make_env, blobs, oracle_env, reference_draws


@pytest.mark.parametrize("task", GOLDEN_TASKS)
def test_oracle_reproduces_reference_golden(oracle, bundled_only, task):
    env = make_env(task)
    g = golden(task)
    traj_no, step_no = reference_draws(env)
    oe = oracle_env(oracle, env, env._drawn_model_no)          # carry tasks: the model (weight) drawn at reset
    rows = [oe.reset_to(traj_no, step_no)]
    absorbing = False
    while not absorbing and len(rows) < 1001:
        obs, _, absorbing = oe.step(np.random.randn(env.info.action_space.shape[0]) * 0.1)
        rows.append(obs)
    rows = np.array(rows)
    if task in PINNED_ROWS:
        n = PINNED_ROWS[task]
        assert np.allclose(rows[:n], g[:n], atol=GOLDEN_ATOL.get(task, 1e-8)), "max abs err %.3e" % np.abs(rows[:n] - g[:n]).max()
        return
    assert rows.shape == g.shape, "episode length (done-flag timing) differs from the golden"
    assert np.allclose(rows, g, atol=GOLDEN_ATOL.get(task, 1e-8)), "max abs err %.3e" % np.abs(rows - g).max()
    # the reset row is pure table lookup: must be (near) bit-exact
    assert np.abs(rows[0] - g[0]).max() < 1e-13
    # terminal row satisfies has_fallen, earlier rows do not
    assert env._has_fallen(rows[-1]) and not any(env._has_fallen(r) for r in rows[:-1])


@pytest.mark.parametrize("task", ["walk", "run"])
def test_oracle_reproduces_4ages_all_goldens(oracle, bundled_only, task):
    """HumanoidTorque4Ages mode "all" (four humanoids in one env): draws of the reference's reset - model (base.py:187-191),
    trajectory within the model's range (base_humanoid_4_ages.py:132-136), sample - then the drawn humanoid's oracle.
    Both are reproduced to 1e-13 over the whole episode. run.all is THE golden that reaches `mjc_BoxBox`: in row 9 the infant's
    two foot boxes pass each other edge to edge, 0.17 mm apart inside the 1 mm margin. That branch of MuJoCo's box-box routine
    (one contact at the midpoint of the closest points of the two edges) is restated exactly (`box_box_edge`); with the general
    convex routine alone (MPR: dist and normal right to 1e-7, contact POSITION up to 1.6 mm off) the row was 5.8e-3 off."""
    g = golden("HumanoidTorque4Ages.%s.all" % task)
    np.random.seed(0)
    model_no = np.random.randint(0, 4)
    np.random.randint(model_no, model_no + 1)
    env = make_env("HumanoidTorque4Ages.%s.%d" % (task, model_no + 1))
    step_no = np.random.randint(0, env.trajectories.trajectory_length)
    oe = oracle_env(oracle, env, 0)
    rows = [oe.reset_to(0, step_no)]
    absorbing = False
    while not absorbing and len(rows) < 1001:
        obs, _, absorbing = oe.step(np.random.randn(env.info.action_space.shape[0]) * 0.1)
        rows.append(obs)
    rows = np.array(rows)
    assert rows.shape == g.shape, "episode length (done-flag timing) differs from the golden"
    assert np.allclose(rows, g), "max abs err %.3e" % np.abs(rows - g).max()


def test_oracle_rollout_threads_deterministic(oracle, bundled_only):
    env = make_env("UnitreeA1.simple")
    mb, tb = blobs(env)
    n1, r1 = oracle.rollout(mb, tb, n_envs=8, n_steps=30, nthreads=1, seed=3)
    n2, r2 = oracle.rollout(mb, tb, n_envs=8, n_steps=30, nthreads=4, seed=3)
    assert n1 == n2 == 240 and r1 == r2 and r1 > 0
