"""
Host (numpy) versions of the reward functors (reference: utils/reward.py:34-117). The CUDA engine fuses the
built-in ones into the step kernel (include/locosim_task.h reward types); these classes keep the reference's
Python surface (`env.reward(state, action, next_state, absorbing)`, CustomReward callbacks) and are what the
parity tests compare the kernel against.
"""
import numpy as np


class RewardInterface:
    def __call__(self, state, action, next_state, absorbing):
        raise NotImplementedError

    def reset_state(self):
        pass


class NoReward(RewardInterface):
    def __call__(self, state, action, next_state, absorbing):
        return 0


class PosReward(RewardInterface):
    def __init__(self, pos_idx):
        self._pos_idx = pos_idx

    def __call__(self, state, action, next_state, absorbing):
        return state[self._pos_idx]


class CustomReward(RewardInterface):
    def __init__(self, reward_callback=None):
        self._reward_callback = reward_callback

    def __call__(self, state, action, next_state, absorbing):
        if self._reward_callback is None:
            return 0
        return self._reward_callback(state, action, next_state)


class TargetVelocityReward(RewardInterface):
    def __init__(self, target_velocity, x_vel_idx):
        self._target_vel = target_velocity
        self._x_vel_idx = x_vel_idx

    def __call__(self, state, action, next_state, absorbing):
        return np.exp(-np.square(state[self._x_vel_idx] - self._target_vel))


class MultiTargetVelocityReward(RewardInterface):
    def __init__(self, target_velocity, x_vel_idx, env_id_len, scalings):
        self._target_vel = target_velocity
        self._env_id_len = env_id_len
        self._scalings = scalings
        self._x_vel_idx = x_vel_idx

    def __call__(self, state, action, next_state, absorbing):
        env_id = state[-self._env_id_len:]
        ind = (np.packbits(env_id.astype(int), bitorder='big') >> (8 - env_id.shape[0]))[0]
        target = self._target_vel * self._scalings[ind]
        return np.exp(-np.square(state[self._x_vel_idx] - target))


class VelocityVectorReward(RewardInterface):
    def __init__(self, x_vel_idx, y_vel_idx, angle_idx, goal_vel_idx):
        self._x_vel_idx, self._y_vel_idx = x_vel_idx, y_vel_idx
        self._angle_idx, self._goal_vel_idx = angle_idx, goal_vel_idx

    def __call__(self, state, action, next_state, absorbing):
        vel = np.array([state[self._x_vel_idx], state[self._y_vel_idx]])
        des = state[self._goal_vel_idx] * state[self._angle_idx]
        return np.exp(-5.0 * np.linalg.norm(vel - des))
