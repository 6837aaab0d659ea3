"""
Host-side trajectory handling (cold path: runs once at env construction, result is uploaded to HBM as the
reset table). Behavioural mirror of /root/reference/loco_mujoco/utils/trajectory.py:8-418 (constructor
arguments, `reset_trajectory`, `create_dataset`, `get_current_sample/get_next_sample`, `get_from_sample`),
written from the behaviour, not the text.
"""
import warnings
from copy import deepcopy

import numpy as np
from scipy import interpolate


class Trajectory:
    """
    Per-observation-key arrays of shape (n_traj, n_samples[, dim]) resampled from `traj_dt` to `control_dt`
    with a cubic spline (reference: trajectory.py:184-234), plus the sampling/flattening helpers the envs use.
    """

    def __init__(self, keys, low, high, joint_pos_idx, interpolate_map, interpolate_remap,
                 traj_path=None, traj_files=None, interpolate_map_params=None, interpolate_remap_params=None,
                 traj_dt=0.002, control_dt=0.01, ignore_keys=None, clip_trajectory_to_joint_ranges=False,
                 traj_info=None, warn=True):
        if (traj_path is None) == (traj_files is None):
            raise AssertionError("Please specify either traj_path or traj_files, but not both.")
        files = np.load(traj_path, allow_pickle=True) if traj_path is not None else traj_files
        self._trajectory_files = {k: np.asarray(v) for k, v in files.items()}

        self.check_if_trajectory_is_in_range(low, high, keys, joint_pos_idx, warn, clip_trajectory_to_joint_ranges)

        keys = list(keys)
        keys += [k for k in self._trajectory_files if k.startswith("goal") and k not in keys]
        if ignore_keys is not None:
            for ik in ignore_keys:
                keys.remove(ik)
        self.keys = keys

        if "split_points" in self._trajectory_files:
            self.split_points = np.asarray(self._trajectory_files["split_points"])
        else:
            n = len(next(iter(self._trajectory_files.values())))
            self.split_points = np.array([0, n])

        self.trajectories = self._extract_trajectory_from_files()
        if traj_info is not None and len(traj_info) != self.number_of_trajectories:
            raise AssertionError("The number of trajectory infos/labels need to be equal to the number of "
                                 "trajectories.")
        self._traj_info = traj_info
        self.traj_dt = traj_dt
        self.control_dt = control_dt
        if self.traj_dt != control_dt:
            self._interpolate_trajectories(interpolate_map, interpolate_remap, interpolate_map_params,
                                           interpolate_remap_params)
        self.subtraj_step_no = 0
        self.traj_no = 0
        self.subtraj = self._get_subtraj(self.traj_no)

    # ------------------------------------------------------------------------------------------------
    def _extract_trajectory_from_files(self):
        out = []
        lengths = {len(self._trajectory_files[k]) for k in self.keys}
        if len(lengths) != 1:
            raise AssertionError("Some observations have different lengths than others. Trajectory is corrupted. ")
        for k in self.keys:
            parts = np.split(self._trajectory_files[k], self.split_points[1:-1])
            if len({len(p) for p in parts}) != 1:
                raise AssertionError("Only trajectories of equal length are currently supported.")
            out.append(np.array(parts))
        return out

    def _interpolate_trajectories(self, map_funct, re_map_funct, map_params, re_map_params):
        assert (map_funct is None) == (re_map_funct is None)
        T = self.trajectory_length
        x = np.arange(T)
        n_new = round(T * (self.traj_dt / self.control_dt))
        x_new = np.linspace(0, T - 1, n_new, endpoint=True)
        per_traj = []
        for i in range(self.number_of_trajectories):
            traj = [obs[i] for obs in self.trajectories]
            if map_funct is not None:
                traj = map_funct(traj) if map_params is None else map_funct(traj, **map_params)
            new = interpolate.interp1d(x, traj, kind="cubic", axis=1)(x_new)
            if re_map_funct is not None:
                new = re_map_funct(new) if re_map_params is None else re_map_funct(new, **re_map_params)
            per_traj.append(new)
        self.trajectories = [np.array([t[k] for t in per_traj]) for k in range(len(self.keys))]
        self.split_points = np.concatenate([[0], np.cumsum([len(self.trajectories[0][k])
                                                            for k in range(self.number_of_trajectories)])])

    # ------------------------------------------------------------------------------------------------
    def create_dataset(self, ignore_keys=None, state_callback=None, state_callback_params=None):
        """states / next_states / absorbing / last (reference: trajectory.py:104-151)."""
        flat = dict(zip(self.keys, deepcopy(self.flattened_trajectories())))
        if ignore_keys is not None:
            for k in ignore_keys:
                del flat[k]
        states = np.concatenate(list(flat.values()), axis=1)
        if state_callback is not None:
            states = np.array([state_callback(s, **state_callback_params) for s in states])
        chunks = np.split(states, self.split_points[1:-1])
        out = dict(states=np.concatenate([c[:-1] for c in chunks]),
                   next_states=np.concatenate([c[1:] for c in chunks]))
        out["absorbing"] = np.zeros(len(out["states"]))
        out["last"] = np.concatenate([np.concatenate([np.zeros(len(c) - 2), [1.0]]) for c in chunks])
        if self._traj_info is not None:
            out["info"] = np.array([[l] * self.trajectory_length for l in self._traj_info]).reshape(-1)
        return out

    def reset_trajectory(self, substep_no=None, traj_no=None):
        """Random (legacy global numpy RNG, like the reference) or explicit (traj, step) sample."""
        if traj_no is None:
            self.traj_no = np.random.randint(0, self.number_of_trajectories)
        else:
            assert 0 <= traj_no <= self.number_of_trajectories
            self.traj_no = traj_no
        if substep_no is None:
            self.subtraj_step_no = np.random.randint(0, self.trajectory_length)
        else:
            assert 0 <= substep_no <= self.trajectory_length
            self.subtraj_step_no = substep_no
        self.subtraj = self._get_subtraj(self.traj_no)
        # recentre the two leading (root x / y) entries on the chosen sample
        self.subtraj[0] -= self.subtraj[0][self.subtraj_step_no]
        self.subtraj[1] -= self.subtraj[1][self.subtraj_step_no]
        return [obs[self.subtraj_step_no] for obs in self.subtraj]

    def check_if_trajectory_is_in_range(self, low, high, keys, j_idx, warn, clip_trajectory_to_joint_ranges):
        if not (warn or clip_trajectory_to_joint_ranges):
            return
        j_idx = j_idx[2:]
        highs = dict(zip(keys[2:], high))
        lows = dict(zip(keys[2:], low))
        for i, (k, d) in enumerate(self._trajectory_files.items()):
            if i in j_idx and k in keys:
                if warn:
                    clip_message = "Clipping the trajectory into range!" if clip_trajectory_to_joint_ranges else ""
                    if np.max(d) > highs[k]:
                        warnings.warn("Trajectory violates joint range in %s. Maximum in trajectory is %f "
                                      "and maximum range is %f. %s" % (k, np.max(d), highs[k], clip_message),
                                      RuntimeWarning)
                    elif np.min(d) < lows[k]:
                        warnings.warn("Trajectory violates joint range in %s. Minimum in trajectory is %f "
                                      "and minimum range is %f. %s" % (k, np.min(d), lows[k], clip_message),
                                      RuntimeWarning)
                if clip_trajectory_to_joint_ranges:
                    self._trajectory_files[k] = np.clip(self._trajectory_files[k], lows[k], highs[k])

    def get_current_sample(self):
        return self._get_ith_sample_from_subtraj(self.subtraj_step_no)

    def get_next_sample(self):
        self.subtraj_step_no += 1
        if self.subtraj_step_no == self.trajectory_length:
            return None
        return self._get_ith_sample_from_subtraj(self.subtraj_step_no)

    def get_from_sample(self, sample, key):
        assert len(sample) == len(self.keys)
        return sample[self.get_idx(key)]

    def get_idx(self, key):
        return self.keys.index(key)

    def flattened_trajectories(self):
        out = []
        for obs in self.trajectories:
            if obs.ndim == 2:
                out.append(obs.reshape((-1, 1)))
            elif obs.ndim == 3:
                out.append(obs.reshape((-1, obs.shape[2])))
            else:
                raise ValueError("Unsupported shape of observation %s." % (obs.shape,))
        return out

    def _get_subtraj(self, i):
        return [obs[i].copy() for obs in self.trajectories]

    def _get_ith_sample_from_subtraj(self, i):
        return [np.array(obs[i].copy()).flatten() for obs in self.subtraj]

    @property
    def number_obs_trajectory(self):
        return len(self.trajectories)

    @property
    def trajectory_length(self):
        return self.trajectories[0].shape[1]

    @property
    def number_of_trajectories(self):
        return self.trajectories[0].shape[0]
