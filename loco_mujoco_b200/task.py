"""
TaskSpec: everything of LocoEnv.step()/reset() that is not mj_step, flattened for the engines
(wire format: include/locosim_task.h).
"""
import numpy as np

MAGIC = 0x5441534B
VERSION = 4
OBS_QPOS, OBS_QVEL, OBS_GOAL, OBS_GRF, OBS_PARAM = 0, 1, 2, 3, 4
GRF_FLOOR = 127
REWARD_NONE, REWARD_TARGET_VELOCITY, REWARD_VELOCITY_VECTOR, REWARD_POS, REWARD_TRACKING = 0, 1, 2, 3, 4


class TaskSpec:
    def __init__(self, obs_src_type, obs_src_idx, done_terms, reward_type, reward_ints, reward_params, act_mean,
                 act_delta, n_substeps, table, n_goal, recenter, use_absorbing=True, act_idx=None, n_grf=0,
                 grf_group=None, random_rot=None, tracking=None):
        """
        obs_src_type/idx : per observation entry, where it is gathered from (qpos / qvel / per-episode goal feature)
        done_terms       : list of (obs_index, lo, hi); fallen if obs < lo or obs > hi (strict, like the reference)
        table            : float64 [n_traj, traj_len, nq + nv + n_goal] reset table (qpos, qvel, goal features)
        recenter         : the two qpos indices zeroed at reset (root x / y; trajectory.py:268-269)
        n_grf, grf_group : use_foot_forces: number of foot groups and the per-geom group id (-1 / k / GRF_FLOOR)
        random_rot       : None or (qpos index of the yaw joint, dof index of root vx, dof index of root vy): setup_random_rot
        tracking         : (w_pose, k_pose, w_vel, k_vel) of REWARD_TRACKING (include/locosim_task.h)
        """
        self.obs_src_type = np.asarray(obs_src_type, dtype=np.int32)
        self.obs_src_idx = np.asarray(obs_src_idx, dtype=np.int32)
        self.done_terms = list(done_terms)
        self.reward_type = int(reward_type)
        self.reward_ints = list(reward_ints) + [0] * (4 - len(reward_ints))
        self.reward_params = list(reward_params) + [0.0] * (2 - len(reward_params))
        self.act_mean = np.asarray(act_mean, dtype=np.float64)
        self.act_delta = np.asarray(act_delta, dtype=np.float64)
        self.n_substeps = int(n_substeps)
        self.table = np.ascontiguousarray(table, dtype=np.float64)
        self.n_goal = int(n_goal)
        self.recenter = list(recenter)
        self.use_absorbing = bool(use_absorbing)
        self.act_idx = np.arange(len(self.act_mean), dtype=np.int32) if act_idx is None else \
            np.asarray(act_idx, dtype=np.int32)
        self.n_grf = int(n_grf)
        self.grf_group = np.zeros(0, dtype=np.int32) if grf_group is None else np.asarray(grf_group, dtype=np.int32)
        self.random_rot = [-1, -1, -1] if random_rot is None else [int(x) for x in random_rot]
        self.tracking = [0.0, 0.0, 0.0, 0.0] if tracking is None else [float(x) for x in tracking]

    @property
    def obs_dim(self):
        return len(self.obs_src_type)

    def pack(self):
        n_traj, traj_len, _ = self.table.shape
        ih = np.zeros(24, dtype=np.int32)
        ih[16:18] = [self.n_grf, len(self.grf_group)]
        ih[18:21] = self.random_rot
        ih[:16] = [MAGIC, VERSION, self.obs_dim, len(self.done_terms), self.reward_type, self.n_substeps, n_traj,
                   traj_len, self.n_goal, self.recenter[0], self.recenter[1]] + self.reward_ints + [int(self.use_absorbing)]
        ints = np.concatenate([ih, self.obs_src_type, self.obs_src_idx,
                               np.array([t[0] for t in self.done_terms], dtype=np.int32), self.act_idx,
                               self.grf_group]).astype(np.int32)
        rh = np.zeros(8, dtype=np.float64)
        rh[:2] = self.reward_params
        rh[2:6] = self.tracking
        reals = np.concatenate([rh, self.act_mean, self.act_delta,
                                np.array([t[1] for t in self.done_terms], dtype=np.float64),
                                np.array([t[2] for t in self.done_terms], dtype=np.float64),
                                self.table.ravel()])
        return ints, reals
