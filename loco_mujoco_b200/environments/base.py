"""
LocoEnv facade over the batched CUDA engine.

Same public surface as the reference's `LocoEnv` (/root/reference/loco_mujoco/environments/base.py:25-969):
`make`, `register`, `get_all_task_names`, `list_registered_loco_mujoco`, `reset`, `step`, `reward`, `is_absorbing`,
`create_dataset`, `play_trajectory`, `load_trajectory`, `get_obs_idx`, `get_all_observation_keys`,
`get_kinematic_obs_mask`, `info.observation_space / action_space`, `dt`; plus the additive kwargs
`num_envs`, `device`, `seed`:

* `num_envs=None` (default): drop-in single-env mode. `reset()` returns a float64 numpy observation, `step(a)` returns
  `(obs, reward, absorbing, info)`; trajectory sampling uses the legacy global numpy RNG exactly like the reference,
  so `np.random.seed(0)` reproduces the reference's golden rollouts (to fp32 tolerance). No auto-reset.
* `num_envs=N`: batched mode. `reset()` -> `obs[N, D]` torch.cuda float32, `step(a[N, nu])` ->
  `(obs, reward, absorbing, info)` with `info["next_obs"]` (the observation to act on after the in-kernel auto-reset).

All physics runs in the CUDA engine (loco_mujoco_b200/csrc); there is no CPU path.
"""
import os
import warnings
from copy import deepcopy
from itertools import product

import numpy as np

from .. import mjcf, modelpack
from ..task import (TaskSpec, OBS_QPOS, OBS_QVEL, OBS_GOAL, OBS_GRF, OBS_PARAM, GRF_FLOOR, REWARD_NONE, REWARD_TARGET_VELOCITY,
                    REWARD_POS, REWARD_TRACKING)
from ..trajectory import Trajectory
from ..utils.reward import NoReward, CustomReward, TargetVelocityReward, PosReward


class ObservationType:
    """Subset of mushroom_rl.utils.mujoco.ObservationType used by the in-scope envs."""
    JOINT_POS = "JOINT_POS"
    JOINT_VEL = "JOINT_VEL"
    SITE_ROT = "SITE_ROT"


class Box:
    """Minimal stand-in for mushroom_rl.utils.spaces.Box (low / high / shape)."""

    def __init__(self, low, high):
        self.low = np.array(low, dtype=np.float64)
        self.high = np.array(high, dtype=np.float64)

    @property
    def shape(self):
        return self.low.shape


class MDPInfo:
    def __init__(self, observation_space, action_space, gamma, horizon, dt):
        self.observation_space = observation_space
        self.action_space = action_space
        self.gamma = gamma
        self.horizon = horizon
        self.dt = dt

    @property
    def size(self):
        return self.observation_space.shape + self.action_space.shape

    @property
    def shape(self):
        return self.observation_space.shape + self.action_space.shape


class ObservationHelper:
    """Index bookkeeping of the flat observation (what mushroom's ObservationHelper provides to the reference)."""

    def __init__(self, observation_spec, model):
        self.observation_spec = list(observation_spec)
        self.obs_idx_map = {}
        self.obs_low, self.obs_high = [], []
        self.joint_pos_idx, self.joint_vel_idx = [], []
        i = 0
        for key, name, ot in self.observation_spec:
            n = 9 if ot == ObservationType.SITE_ROT else 1
            self.obs_idx_map[key] = list(range(i, i + n))
            if ot == ObservationType.JOINT_POS:
                j = model.joint_id(name)
                if model.jnt_limited[j]:
                    self.obs_low.append(model.jnt_range[j, 0]); self.obs_high.append(model.jnt_range[j, 1])
                else:
                    self.obs_low.append(-np.inf); self.obs_high.append(np.inf)
                self.joint_pos_idx.append(i)
            else:
                self.obs_low += [-np.inf] * n
                self.obs_high += [np.inf] * n
                if ot == ObservationType.JOINT_VEL:
                    self.joint_vel_idx.append(i)
            i += n
        self.obs_length = i

    def get_all_observation_keys(self):
        return [k for k, _, _ in self.observation_spec]

    def get_from_obs(self, obs, key):
        return obs[self.obs_idx_map[key]]

    def get_joint_pos_from_obs(self, obs):
        return obs[self.joint_pos_idx]

    def get_joint_vel_from_obs(self, obs):
        return obs[self.joint_vel_idx]


def reference_data_root():
    """Directory that holds `environments/data` and `datasets` of a loco_mujoco checkout, or None."""
    if os.environ.get("LOCO_MUJOCO_B200_FORCE_BUNDLED"):
        return None
    cands = [os.environ.get("LOCO_MUJOCO_PATH")]
    try:
        import importlib.util
        spec = importlib.util.find_spec("loco_mujoco")
        if spec is not None and spec.origin:
            cands.append(os.path.dirname(spec.origin))
    except Exception:
        pass
    cands.append("/root/reference/loco_mujoco")
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "environments", "data")):
            return c
    return None


ASSET_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets")


class LocoEnv:
    """Base class of all batched locomotion environments."""

    _registered_envs = dict()

    def __init__(self, xml_handles, action_spec, observation_spec, collision_groups=None, gamma=0.99, horizon=1000,
                 n_substeps=10, reward_type=None, reward_params=None, traj_params=None, random_start=True,
                 init_step_no=None, timestep=0.001, use_foot_forces=False, default_camera_mode="follow",
                 use_absorbing_states=True, domain_randomization_config=None, parallel_dom_rand=True,
                 N_worker_per_xml_dom_rand=4, num_envs=None, device="cuda:0", seed=0, env_id_offset=0,
                 compiled_model=None, copy_outputs=True, convex_collisions=True, warps_per_block=None, **viewer_params):
        if type(xml_handles) != list:
            xml_handles = [xml_handles]
        self._xml_handles = xml_handles
        # several handles = a multi-model env (carry tasks: one model per weight; base.py:86-88,183-191). All models share
        # topology and sizes, so the engine runs them as rows of the parameter pool (one row per model).
        # use_foot_forces (base.py:93-98 of the reference): n_intermediate_steps x mj_step(1) with a contact-force hook
        # after each one. The engine always runs n_substeps physics steps per control step; dt is unchanged.
        self._collision_groups = list(collision_groups) if collision_groups else []
        self._domain_rand_config = domain_randomization_config
        self._domain_rand_pool_size = viewer_params.pop("domain_randomization_pool_size", 64)
        self._domain_rand = None
        self._timestep = timestep
        self._n_substeps = n_substeps
        self._n_intermediate_steps = 1
        self._use_foot_forces = use_foot_forces
        if compiled_model is not None:
            self._models = list(compiled_model) if isinstance(compiled_model, (list, tuple)) else [compiled_model]
        else:
            self._models = [mjcf.compile_model(h, timestep=timestep) for h in xml_handles]
        self._model = self._models[0]
        if timestep is None:
            self._timestep = self._model.opt_timestep
        self._model_user_features = [()] * len(self._models)     # per model: values observable through OBS_PARAM
        if len(self._models) > 1 and domain_randomization_config is not None:
            raise NotImplementedError("domain randomisation of a multi-model env")
        self._action_spec = list(action_spec) if len(action_spec) else list(self._model.actuator_names)
        self._action_indices = [self._model.actuator_id(n) for n in self._action_spec]
        if sorted(self._action_indices) != list(range(self._model.nu)):
            raise NotImplementedError("action_spec must cover all actuators of the (modified) model")
        self.obs_helper = ObservationHelper(observation_spec, self._model)
        self.obs_helpers = [self.obs_helper]
        lo = self._model.actuator_ctrlrange[self._action_indices, 0].copy()
        hi = self._model.actuator_ctrlrange[self._action_indices, 1].copy()
        self.info = MDPInfo(Box(self.obs_helper.obs_low, self.obs_helper.obs_high), Box(lo, hi), gamma, horizon, self.dt)

        self._reward_type, self._reward_params = reward_type, reward_params
        self._reward_function = self._get_reward_function(reward_type, reward_params)
        self.info.observation_space = Box(*self._get_observation_space())
        low, high = self.info.action_space.low.copy(), self.info.action_space.high.copy()
        self.norm_act_mean = (high + low) / 2.0
        self.norm_act_delta = (high - low) / 2.0
        self.info.action_space.low[:] = -1.0
        self.info.action_space.high[:] = 1.0

        self._dataset = None
        self.trajectories = None
        if traj_params:
            self.load_trajectory(traj_params)
        self._random_start = random_start
        self._init_step_no = init_step_no
        self._use_absorbing_states = use_absorbing_states

        self._num_envs = num_envs
        self._device = device
        self._seed = seed
        self._env_id_offset = env_id_offset
        # Batched mode: the engine writes every step into the same device buffers. copy_outputs=True (default) hands the
        # caller private copies (two device-to-device copies per step), so that appending them to a rollout buffer is safe;
        # copy_outputs=False returns views that are only valid until the next step()/reset() (zero-copy, benchmark use).
        self._copy_outputs = bool(copy_outputs)
        # convex_collisions=False drops the mesh-mesh / box-mesh candidate pairs (bone against bone: mjc_Convex / MPR) from the
        # engine's pair table: faster (HumanoidTorque: ~3x), but not what the reference simulates. Default: on.
        self._convex_collisions = bool(convex_collisions)
        # launch geometry of the step kernel (envs per block); None = the engine's choice. Scheduling only (MixedBatch uses it)
        self._warps_per_block = warps_per_block
        self._engine = None
        self._obs = None

    # ---------------------------------------------------------------------------------------------------
    # construction helpers
    # ---------------------------------------------------------------------------------------------------
    @property
    def dt(self):
        return self._timestep * self._n_intermediate_steps * self._n_substeps

    @property
    def num_envs(self):
        return 1 if self._num_envs is None else self._num_envs

    @property
    def batched(self):
        return self._num_envs is not None

    def load_trajectory(self, traj_params, warn=True):
        if self.trajectories is not None:
            warnings.warn("New trajectories loaded, which overrides the old ones.", RuntimeWarning)
        if "processed" in traj_params:
            self.trajectories = _ProcessedTrajectory(traj_params["processed"])
        else:
            self.trajectories = Trajectory(keys=self.get_all_observation_keys(),
                                           low=self.info.observation_space.low,
                                           high=self.info.observation_space.high,
                                           joint_pos_idx=self.obs_helper.joint_pos_idx,
                                           interpolate_map=self._interpolate_map,
                                           interpolate_remap=self._interpolate_remap,
                                           interpolate_map_params=self._get_interpolate_map_params(),
                                           interpolate_remap_params=self._get_interpolate_remap_params(),
                                           warn=warn, **traj_params)
        self._engine = None

    def get_all_observation_keys(self):
        return self.obs_helper.get_all_observation_keys()

    # ---------------------------------------------------------------------------------------------------
    # TaskSpec: flatten obs / done / reward / reset table for the engine
    # ---------------------------------------------------------------------------------------------------
    def _has_fallen_terms(self):
        """list of (observation key, low, high): fallen if value < low or value > high."""
        raise NotImplementedError

    def _goal_features(self, sample):
        """Per-episode constant observation features derived from a trajectory sample (goal-conditioned envs)."""
        return []

    def _n_goal(self):
        return 0

    def _obs_sources(self):
        """(src_type, src_idx) for every entry of the final observation vector."""
        types, idxs = [], []
        for key, name, ot in self.obs_helper.observation_spec[2:]:
            if ot == ObservationType.JOINT_POS:
                types.append(OBS_QPOS); idxs.append(self._model.joint_id(name))
            elif ot == ObservationType.JOINT_VEL:
                types.append(OBS_QVEL); idxs.append(self._model.joint_id(name))
            else:
                raise NotImplementedError(ot)
        return types, idxs

    def _reward_spec(self):
        rt = self._reward_type
        if rt is None:
            return REWARD_NONE, [], []
        if rt == "target_velocity":
            return REWARD_TARGET_VELOCITY, [self.get_obs_idx("dq_pelvis_tx")[0]], [self._reward_params["target_velocity"]]
        if rt == "x_pos":
            return REWARD_POS, [self.get_obs_idx("q_pelvis_tx")[0]], []
        if rt == "custom":
            return REWARD_NONE, [], []      # evaluated on the host from the returned tensors
        if rt == "tracking":
            return REWARD_TRACKING, [], []  # weights travel in TaskSpec.tracking
        raise NotImplementedError(rt)

    def _reset_table(self):
        """[n_traj, T, nq + nv + n_goal] from the (interpolated) trajectories."""
        tr = self.trajectories
        nq = self._model.nq
        n_traj, T = tr.number_of_trajectories, tr.trajectory_length
        table = np.zeros((n_traj, T, 2 * nq + self._n_goal()))
        spec = self.obs_helper.observation_spec
        for k, (key, name, ot) in enumerate(spec):
            arr = tr.trajectories[tr.keys.index(key)]
            if ot == ObservationType.JOINT_POS:
                table[:, :, self._model.joint_id(name)] = arr
            elif ot == ObservationType.JOINT_VEL:
                table[:, :, nq + self._model.joint_id(name)] = arr
        if self._n_goal():
            for i in range(n_traj):
                for j in range(T):
                    sample = [obs[i][j] for obs in tr.trajectories]
                    table[i, j, 2 * nq:] = self._goal_features(sample)
        return table

    def task_spec(self):
        types, idxs = self._obs_sources()
        keys_to_idx = {}
        done_terms = []
        for key, lo, hi in self._has_fallen_terms():
            done_terms.append((self.get_obs_idx(key)[0], lo, hi))
        rtype, rints, rparams = self._reward_spec()
        spec = self.obs_helper.observation_spec
        recenter = [self._model.joint_id(spec[0][1]), self._model.joint_id(spec[1][1])]
        if self.trajectories is None:
            raise ValueError("the CUDA engine needs trajectory data for resets (pass traj_params)")
        n_grf, grf_group = self._grf_spec()
        types = list(types) + [OBS_GRF] * (3 * n_grf)           # mean ground forces, then per-model features (weight)
        idxs = list(idxs) + list(range(3 * n_grf))
        n_user = len(self._model_user_features[0])
        types += [OBS_PARAM] * n_user
        idxs += list(range(n_user))
        return TaskSpec(types, idxs, done_terms, rtype, rints, rparams, self.norm_act_mean, self.norm_act_delta,
                        self._n_substeps, self._reset_table(), self._n_goal(), recenter, self._use_absorbing_states,
                        act_idx=self._action_indices, n_grf=n_grf, grf_group=grf_group, random_rot=self._random_rot_spec(),
                        tracking=self._tracking_params())

    TRACKING_DEFAULTS = dict(w_pose=0.7, k_pose=2.0, w_vel=0.3, k_vel=0.1)

    def _tracking_params(self):
        """(w_pose, k_pose, w_vel, k_vel) of reward_type="tracking" (this package's mocap-tracking reward, see
        include/locosim_task.h LS_REWARD_TRACKING; the reference has no reward that reads the trajectory)."""
        if self._reward_type != "tracking":
            return None
        prm = dict(self.TRACKING_DEFAULTS, **(self._reward_params or {}))
        return [prm["w_pose"], prm["k_pose"], prm["w_vel"], prm["k_vel"]]

    def _random_rot_spec(self):
        """None, or (qpos index of the yaw joint, dof indices of the root x / y velocity): setup_random_rot."""
        return None

    def domain_randomization_pool(self):
        """[K, P] parameter pool: K consecutive randomised recompilations of the model (see domain_randomization.py)."""
        if self._domain_rand is None:
            from ..domain_randomization import DomainRandomizationHandler
            if self._xml_handles[0] is None:
                self._dr_pool = self._bundled_dr_pool()
                self._domain_rand = "bundled"
                return self._dr_pool
            self._domain_rand = DomainRandomizationHandler([self._xml_handles[0].copy()], self._domain_rand_config,
                                                           timestep=self._timestep)
            self._dr_pool = self._domain_rand.build_pool(self._domain_rand_pool_size)
        return self._dr_pool

    def _bundled_dr_pool(self):
        """Without the MJCF sources (bundled assets only, e.g. on the GPU box) the randomised recompilations cannot be
        made; for the reference's shipped YAML configs a pre-built seeded pool is bundled (tools/build_dr_pools.py)."""
        import json
        task_id = getattr(self, "_task_id", None)
        path = os.path.join(ASSET_DIR, "dr_pools", "%s.npz" % task_id)
        if task_id is None or not os.path.exists(path):
            raise ValueError("domain randomisation needs the MJCF source (a loco_mujoco checkout); no pre-built pool is "
                             "bundled for %s" % task_id)
        d = np.load(path, allow_pickle=False)
        meta = json.loads(str(d["meta"]))
        conf = self._domain_rand_config
        ok = (isinstance(conf, str) and os.path.basename(conf) == meta["yaml"]) or conf == meta["config_used"] or \
            conf == meta["config_shipped"]
        if not ok:
            raise ValueError("the bundled pool of %s was built for %s; other configs need the MJCF source"
                             % (task_id, meta["yaml"]))
        return np.asarray(d["pool"], dtype=np.float64)

    def model_pool(self):
        """[n_models, P] parameter pool of a multi-model env (one row per model incl. its user features)."""
        from ..domain_randomization import pool_row
        return np.stack([pool_row(m, u) for m, u in zip(self._models, self._model_user_features)])

    def _get_engine(self):
        if self._engine is None:
            from ..engine import CudaEngine
            import torch
            dev = torch.device(self._device)
            self._engine = CudaEngine(modelpack.pack(self._model, convex_pairs=self._convex_collisions),
                                      self.task_spec().pack(), self.num_envs,
                                      device=dev.index or 0, seed=self._seed, env_id_offset=self._env_id_offset,
                                      warps_per_block=self._warps_per_block)
            if self._domain_rand_config is not None:
                self._engine.set_param_pool(self.domain_randomization_pool())
            elif len(self._models) > 1 or len(self._model_user_features[0]):
                self._engine.set_param_pool(self.model_pool())
        return self._engine

    # ---------------------------------------------------------------------------------------------------
    # the hot path
    # ---------------------------------------------------------------------------------------------------
    def reward(self, state, action, next_state, absorbing):
        return self._reward_function(state, action, next_state, absorbing)

    def _state_from_obs(self, obs):
        """Observation(s) -> (qpos, qvel) [n, nq] (root x / y = 0), goal features [n, n_goal] or None.
        Reference: LocoEnv._init_sim_from_obs + set_sim_state (base.py:478-497,633-654)."""
        obs = np.atleast_2d(np.asarray(obs, dtype=np.float64))
        types, idxs = self._obs_sources()
        assert obs.shape[1] >= len(types), "observation shorter than the observation specification"
        nq = self._model.nq
        q, v = np.zeros((len(obs), nq)), np.zeros((len(obs), nq))
        goal = np.zeros((len(obs), 4))
        for k, (t, i) in enumerate(zip(types, idxs)):
            if t == OBS_QPOS:
                q[:, i] = obs[:, k]
            elif t == OBS_QVEL:
                v[:, i] = obs[:, k]
            elif t == OBS_GOAL:
                goal[:, i] = obs[:, k]
        return q, v, (goal if self._n_goal() else None)

    def _reset_from_obs(self, eng, obs):
        import torch
        if torch.is_tensor(obs):
            obs = obs.detach().cpu().numpy()
        q, v, goal = self._state_from_obs(obs)
        assert len(q) == self.num_envs, "reset(obs): one observation per env"
        if len(self._models) > 1:
            raise NotImplementedError("reset(obs=...) of a multi-model env")
        eng.reset()                                   # episode bookkeeping (counters, parameter-pool row, foot forces)
        f = lambda a: torch.tensor(a, dtype=torch.float32, device=eng.device).contiguous()
        eng.set_state(f(q), f(v))
        if goal is not None:
            eng.set_goal(f(goal))
        out = f(np.atleast_2d(np.asarray(obs, dtype=np.float64))[:, :eng.obs_dim])
        if out.shape[1] < eng.obs_dim:                # foot forces / per-model features: zero / unchanged after a reset
            out = torch.cat([out, eng.next_obs[:, out.shape[1]:]], dim=1)
        eng.next_obs.copy_(out)
        return eng.next_obs

    def reset(self, obs=None):
        eng = self._get_engine()
        import torch
        self._reward_function.reset_state()
        if obs is None:
            # argument checks of the reference's setup() (base.py:220-225)
            if self.trajectories is None and self._random_start:
                raise ValueError("Random start not possible without trajectory data.")
            if self.trajectories is None and self._init_step_no is not None:
                raise ValueError("Setting an initial step is not possible without trajectory data.")
            if self._init_step_no is not None and self._random_start:
                raise ValueError("Either use a random start or set an initial step, not both.")
        if not self.batched:
            # reference semantics incl. the legacy global numpy RNG draw order (base.py:187-191, trajectory.py:253-259)
            model_no = np.random.randint(0, len(self._models))
            self._current_model_idx = model_no
            if obs is not None:
                out = self._reset_from_obs(eng, obs)
                self._obs = out[0].double().cpu().numpy()
                return self._obs.copy()
            if self._random_start:
                sample = self.trajectories.reset_trajectory()
            elif self._init_step_no is not None:
                T, n_traj = self.trajectories.trajectory_length, self.trajectories.number_of_trajectories
                assert self._init_step_no <= T * n_traj
                sample = self.trajectories.reset_trajectory(int(self._init_step_no % T), int(self._init_step_no / T))
            else:
                sample = self.trajectories.reset_trajectory(substep_no=0)
            traj_no, step_no = self.trajectories.traj_no, self.trajectories.subtraj_step_no
            t = torch.tensor([traj_no], dtype=torch.int32, device=eng.device)
            s = torch.tensor([step_no], dtype=torch.int32, device=eng.device)
            r = torch.tensor([model_no], dtype=torch.int32, device=eng.device) if len(self._models) > 1 else None
            ang = self._reset_rotation_angle()
            ang = None if ang is None else torch.tensor([ang], dtype=torch.float32, device=eng.device)
            out = eng.reset(traj_no=t, step_no=s, pool_row=r, rot_angle=ang)
            self._obs = out[0].double().cpu().numpy()
            return self._obs.copy()
        if obs is not None:
            out = self._reset_from_obs(eng, obs)
        else:
            out = eng.reset()
        self._obs = out.clone() if self._copy_outputs else out
        return self._obs

    def set_launch_geometry(self, warps_per_block):
        """Envs (= warps) per block of the step kernel: None = the engine's choice, -1 = always the largest block, k = k.
        Scheduling only. The engine is rebuilt (state lost): call before reset()."""
        self._warps_per_block = warps_per_block
        if getattr(self, "_engine", None) is not None:
            self._engine.close() if hasattr(self._engine, "close") else None
            self._engine = None

    def _reset_rotation_angle(self):
        """Drop-in single-env reset: host-drawn rotation angle of setup_random_rot envs (None: env has no such option)."""
        return None

    def step(self, action):
        eng = self._get_engine()
        import torch
        if not self.batched:
            a = torch.as_tensor(np.asarray(action, dtype=np.float32).reshape(1, -1), device=eng.device)
            obs, reward, done, _ = eng.step(a.contiguous(), auto_reset=False, want_next_obs=False)
            cur = obs[0].double().cpu().numpy()
            absorbing = bool(done[0].item())
            if self._reward_type == "custom":
                r = self._reward_function(self._obs, np.asarray(action), cur, absorbing)
            else:
                r = float(reward[0].item())
            self._obs = cur
            return cur.copy(), r, absorbing, {}
        if action.dtype != torch.float32:
            action = action.float()
        prev = self._obs
        if self._reward_type == "custom" and not self._copy_outputs and prev is not None:
            prev = prev.clone()          # the engine is about to overwrite the buffer `prev` is a view of
        obs, reward, done, next_obs = eng.step(action.contiguous(), auto_reset=True)
        if self._copy_outputs:
            packed = eng.packed_out.clone()
            obs, reward, done = eng.views_of(packed)
            next_obs = next_obs.clone()
        info = {"next_obs": next_obs}
        if self._reward_type == "custom":
            # CustomReward(state, action, next_state): `state` is the observation the action was taken in (base.py:170-176)
            cb = self._reward_function._reward_callback
            reward = cb(prev, action, obs) if cb is not None else torch.zeros_like(reward)
        self._obs = next_obs
        return obs, reward, done.view(torch.bool), info      # (0/1 bytes reinterpreted, no kernel)

    def is_absorbing(self, obs):
        return self._has_fallen(obs) if self._use_absorbing_states else False

    def stop(self):
        pass

    def render(self, record=False):
        raise NotImplementedError("rendering is out of scope (headless engine)")

    # ---------------------------------------------------------------------------------------------------
    # observation bookkeeping (reference: base.py:257-276, 729-775)
    # ---------------------------------------------------------------------------------------------------
    def get_kinematic_obs_mask(self):
        return np.arange(len(self.obs_helper.observation_spec) - 2)

    def get_obs_idx(self, key):
        return [i - 2 for i in self.obs_helper.obs_idx_map[key]]

    def _get_idx(self, keys):
        if type(keys) != list:
            keys = [keys]
        return np.concatenate([self.obs_helper.obs_idx_map[k] for k in keys]) - 2

    def _get_from_obs(self, obs, keys):
        obs = np.concatenate([[0.0, 0.0], obs])
        if type(keys) != list:
            keys = [keys]
        return np.concatenate([self.obs_helper.get_from_obs(obs, k) for k in keys])

    def _get_observation_space(self):
        lo, hi = self.info.observation_space.low[2:], self.info.observation_space.high[2:]
        return self._append_grf_space(lo, hi)

    def _append_grf_space(self, lo, hi):
        if self._use_foot_forces:      # base.py:576-580
            n = self._get_grf_size()
            lo, hi = np.concatenate([lo, -np.ones(n) * np.inf]), np.concatenate([hi, np.ones(n) * np.inf])
        return lo, hi

    @staticmethod
    def _get_grf_size():
        return 12

    def _grf_group_names(self):
        """Foot groups in the order of `_get_ground_forces` (base.py:667-679)."""
        return ["foot_r", "front_foot_r", "foot_l", "front_foot_l"]

    def _grf_spec(self):
        """(n_groups, per-geom group id) for the engine; the floor group is matched against every foot group."""
        if not self._use_foot_forces:
            return 0, None
        groups = dict(self._collision_groups)
        names = self._grf_group_names()
        assert 3 * len(names) == self._get_grf_size()
        gid = -np.ones(self._model.ngeom, dtype=np.int32)
        for g in groups["floor"]:
            gid[self._model.geom_id(g)] = GRF_FLOOR
        for k, name in enumerate(names):
            for g in groups[name]:
                gid[self._model.geom_id(g)] = k
        return len(names), gid

    def _create_observation(self, obs):
        return np.concatenate([obs[2:]]).flatten()

    # hidable parts of the observation, in observation order (POMDP masks; base_robot_humanoid.py:38-90,
    # base_humanoid_4_ages.py:187-241)
    _hidable_obs = ("positions", "velocities", "foot_forces")

    def _len_qpos_qvel(self):
        spec = self.obs_helper.observation_spec
        n_pos = sum(1 for _, _, ot in spec if ot == ObservationType.JOINT_POS)
        n_vel = sum(1 for _, _, ot in spec if ot == ObservationType.JOINT_VEL)
        return n_pos, n_vel

    def get_mask(self, obs_to_hide):
        """Boolean mask over the observation: False for the parts named in `obs_to_hide` ("positions", "velocities",
        "foot_forces", and per env "weight" / "env_type")."""
        if type(obs_to_hide) == str:
            obs_to_hide = (obs_to_hide,)
        assert all(x in self._hidable_obs for x in obs_to_hide), "Some of the observations you want to hide are not" \
                                                                 "supported. Valid observations to hide are %s." \
                                                                 % (self._hidable_obs,)
        pos_dim, vel_dim = self._len_qpos_qvel()
        mask = [np.full(pos_dim - 2, "positions" not in obs_to_hide), np.full(vel_dim, "velocities" not in obs_to_hide)]
        if self._use_foot_forces:
            mask.append(np.full(self._get_grf_size(), "foot_forces" not in obs_to_hide))
        else:
            assert "foot_forces" not in obs_to_hide, "Creating a mask to hide foot forces without activating " \
                                                     "the latter is not allowed."
        n_user = len(self._model_user_features[0])
        if n_user:
            key = self._user_feature_name()
            mask.append(np.full(n_user, key not in obs_to_hide))
        else:
            assert not any(k in obs_to_hide for k in ("weight", "env_type")), \
                "Creating a mask to hide the carried weight / env type without activating the latter is not allowed."
        return np.concatenate(mask).ravel().astype(bool)

    def _user_feature_name(self):
        return "weight"

    def _preprocess_action(self, action):
        return (np.asarray(action).copy() * self.norm_act_delta) + self.norm_act_mean

    def _has_fallen(self, obs, return_err_msg=False):
        fallen, msg = False, ""
        for key, lo, hi in self._has_fallen_terms():
            v = obs[self.get_obs_idx(key)[0]]
            if v < lo or v > hi:
                fallen = True
                msg += "%s condition violated (%f not in [%f, %f]).\n" % (key, v, lo, hi)
        return (fallen, msg) if return_err_msg else fallen

    def _get_reward_function(self, reward_type, reward_params):
        if reward_type == "custom":
            return CustomReward(**reward_params)
        if reward_type == "target_velocity":
            idx = self.get_obs_idx("dq_pelvis_tx")
            assert len(idx) == 1
            return TargetVelocityReward(x_vel_idx=idx[0], **reward_params)
        if reward_type == "x_pos":
            idx = self.get_obs_idx("q_pelvis_tx")
            assert len(idx) == 1
            return PosReward(pos_idx=idx[0])
        if reward_type is None or reward_type == "tracking":      # tracking: evaluated in the kernel only (no host functor)
            return NoReward()
        raise NotImplementedError("The specified reward has not been implemented: %s" % reward_type)

    # ---------------------------------------------------------------------------------------------------
    # datasets / replay (reference: base.py:278-386)
    # ---------------------------------------------------------------------------------------------------
    def create_dataset_device(self):
        """`create_dataset()` built ON the device from the reset table that already lives in HBM: dict of torch.cuda float32
        tensors (states / next_states [n, obs_dim] in observation layout, absorbing / last [n]); same content as the host
        `create_dataset()` with the env's default ignore keys (reference: base.py:278-312, utils/trajectory.py:104-151)."""
        import torch
        states, nxt, last = self._get_engine().create_dataset()
        return dict(states=states, next_states=nxt, absorbing=torch.zeros_like(last), last=last)

    def create_dataset(self, ignore_keys=None):
        if self._dataset is None:
            if self.trajectories is None:
                raise ValueError("No trajectory was passed to the environment. To create a dataset pass a trajectory "
                                 "first.")
            dataset = self.trajectories.create_dataset(ignore_keys=ignore_keys)
            for state in dataset["states"]:
                fallen, msg = self._has_fallen(state, return_err_msg=True)
                if fallen:
                    raise ValueError("Some of the states in the created dataset are terminal states. This should not "
                                     "happen.\n\nViolations:\n" + msg)
            self._dataset = deepcopy(dataset)
            return dataset
        return deepcopy(self._dataset)

    def play_trajectory(self, n_episodes=None, n_steps_per_episode=None, render=False, record=False,
                        recorder_params=None):
        """Kinematic replay (no physics is advanced in the reference either: base.py:314-386)."""
        assert self.trajectories is not None
        if render or record:
            raise NotImplementedError("rendering is out of scope (headless engine)")
        big = np.iinfo(np.int32).max
        n_episodes = big if n_episodes is None else n_episodes
        n_steps_per_episode = big if n_steps_per_episode is None else n_steps_per_episode
        self.trajectories.reset_trajectory()
        for _ in range(n_episodes):
            for _ in range(n_steps_per_episode):
                sample = self.trajectories.get_next_sample()
                if sample is None:
                    self.trajectories.reset_trajectory()
                    sample = self.trajectories.get_current_sample()
                obs = self._create_observation(np.concatenate(sample))
                if self._has_fallen(obs):
                    print("Has fallen!")
            self.trajectories.reset_trajectory()

    # ---------------------------------------------------------------------------------------------------
    # interpolation hooks (identity by default, reference: base.py:855-893)
    # ---------------------------------------------------------------------------------------------------
    def _get_interpolate_map_params(self):
        return None

    def _get_interpolate_remap_params(self):
        return None

    @staticmethod
    def _interpolate_map(traj, **interpolate_map_params):
        return np.array(traj)

    @staticmethod
    def _interpolate_remap(traj, **interpolate_remap_params):
        return [obs for obs in traj]

    @staticmethod
    def _delete_from_xml_handle(xml_handle, joints_to_remove, motors_to_remove, equ_constraints):
        return xml_handle.delete(joints_to_remove, motors_to_remove, equ_constraints)

    @property
    def xml_handle(self):
        return self._xml_handles[0]

    @property
    def xml_handles(self):
        return self._xml_handles

    # ---------------------------------------------------------------------------------------------------
    # registry / factory (reference: base.py:820-832, 950-967; mushroom Environment.make)
    # ---------------------------------------------------------------------------------------------------
    @classmethod
    def register(cls):
        LocoEnv._registered_envs[cls.__name__] = cls

    @staticmethod
    def list_registered_loco_mujoco():
        return list(LocoEnv._registered_envs.keys())

    @staticmethod
    def make(env_name, *args, **kwargs):
        """`LocoEnv.make("UnitreeA1.simple.real", num_envs=4096)`: dotted id -> env class .generate(*id_parts)."""
        if "." in env_name:
            parts = env_name.split(".")
            env_name, args = parts[0], tuple(parts[1:]) + tuple(args)
        if env_name not in LocoEnv._registered_envs:
            raise KeyError("unknown environment %r; registered: %s" % (env_name, LocoEnv.list_registered_loco_mujoco()))
        env = LocoEnv._registered_envs[env_name]
        return env.generate(*args, **kwargs) if hasattr(env, "generate") else env(*args, **kwargs)

    @classmethod
    def get_all_task_names(cls):
        names = []
        for e in cls.list_registered_loco_mujoco():
            env = cls._registered_envs[e]
            for conf in env.valid_task_confs.get_all_combinations():
                names.append(".".join([env.__name__] + list(conf.values())))
        return names


class _ProcessedTrajectory(Trajectory):
    """Trajectory restored from already-interpolated arrays (bundled assets; see tools/build_assets.py)."""

    def __init__(self, d):
        self.keys = [str(k) for k in d["keys"]]
        self.trajectories = [np.asarray(d["traj_%d" % i]) for i in range(len(self.keys))]
        self.split_points = np.asarray(d["split_points"])
        self._traj_info = None
        self.traj_dt = self.control_dt = float(d["control_dt"])
        self.subtraj_step_no = 0
        self.traj_no = 0
        self.subtraj = self._get_subtraj(self.traj_no)

    def to_dict(self):
        d = dict(keys=np.array(self.keys), split_points=self.split_points, control_dt=self.control_dt)
        for i, t in enumerate(self.trajectories):
            d["traj_%d" % i] = t
        return d


def processed_trajectory_dict(traj):
    d = dict(keys=np.array(traj.keys), split_points=np.asarray(traj.split_points), control_dt=traj.control_dt)
    for i, t in enumerate(traj.trajectories):
        d["traj_%d" % i] = t
    return d


class ValidTaskConf:
    """All valid task configurations of an environment (reference: base.py:972-1041)."""

    def __init__(self, tasks=None, modes=None, data_types=None, non_combinable=None):
        self.tasks, self.modes, self.data_types, self.non_combinable = tasks, modes, data_types, non_combinable
        if non_combinable is not None:
            for nc in non_combinable:
                assert len(nc) == 3

    def get_all(self):
        return deepcopy(self.tasks), deepcopy(self.modes), deepcopy(self.data_types), deepcopy(self.non_combinable)

    def get_all_combinations(self):
        confs = []
        for t, m, dt in product(self.tasks or [None], self.modes or [None], self.data_types or [None]):
            conf = {}
            if t is not None:
                conf["task"] = t
            if m is not None:
                conf["mode"] = m
            if dt is not None:
                conf["data_type"] = dt
            if self.non_combinable is not None:
                # (reference quirk kept: a conf is appended once per non-combinable rule it does not match)
                for bad_t, bad_m, bad_dt in self.non_combinable:
                    if not ((t == bad_t or bad_t is None) and (m == bad_m or bad_m is None) and
                            (dt == bad_dt or bad_dt is None)):
                        confs.append(conf)
            else:
                confs.append(conf)
        return confs
