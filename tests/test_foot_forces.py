"""use_foot_forces: oracle-level checks on CPU. No reference golden pins this path (it is off in
tests/test_environments.py), so the decode of the constraint forces is pinned physically: at rest the normal forces of
all contacts carry the weight, and the observation entries are exactly the first floor contact of each foot group."""
import ctypes

import numpy as np
import pytest

from helpers import make_env, blobs


@pytest.mark.parametrize("task,steps", [("UnitreeA1.simple", 250), ("Atlas.walk", 800)])
def test_oracle_foot_forces_at_rest(oracle, bundled_only, task, steps):
    lib = oracle.lib
    lib.ref_ncon.argtypes = [ctypes.c_void_p]
    lib.ref_get_contact.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    env = make_env(task, use_foot_forces=True, use_absorbing_states=False)
    spec = env.task_spec()
    n_grf = spec.n_grf
    assert spec.obs_dim == env.info.observation_space.shape[0] and 3 * n_grf == env._get_grf_size()
    oe = oracle.env(*blobs(env))
    obs = oe.reset_to(0, 10)
    assert np.all(obs[-3 * n_grf:] == 0.0)                     # reset observation: empty running mean
    for _ in range(steps):                                     # zero torque: the robot collapses and comes to rest
        obs, _, _ = oe.step(np.zeros(env.info.action_space.shape[0]))
    q, v = oe.get_state()
    assert np.abs(v).max() < 5e-2, "not at rest"
    total, first = 0.0, {}
    for k in range(lib.ref_ncon(oe.sim)):
        out = np.zeros(16)
        lib.ref_get_contact(oe.sim, k, out.ctypes.data_as(ctypes.c_void_p))      # mj_contactForce decode
        total += out[10]
        g1, g2 = spec.grf_group[int(out[7])], spec.grf_group[int(out[8])]
        g = g2 if g1 == 127 else (g1 if g2 == 127 else -1)
        if 0 <= g < 127 and g not in first:
            first[g] = out[10:13].copy()
    weight = env._model.body_mass.sum() * 9.81
    assert abs(total - weight) < 0.01 * weight, (total, weight)
    grf = obs[-3 * n_grf:].reshape(n_grf, 3) * 1000.0
    for g in range(n_grf):
        want = first.get(g, np.zeros(3))
        assert np.allclose(grf[g], want, atol=1e-3 * max(1.0, np.abs(want).max())), (g, grf[g], want)
    assert len(first) > 0
