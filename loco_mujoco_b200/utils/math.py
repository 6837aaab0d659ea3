"""
Angle helpers used by the UnitreeA1 goal features. Behavioural mirror of
/root/reference/loco_mujoco/utils/math.py:5-78 and of the two mushroom_rl.utils.angles functions it imports
(euler_to_mat / mat_to_euler: extrinsic xyz Euler angles <-> rotation matrix, third-party, restated here).
"""
import numpy as np

_EPS4 = np.finfo(np.float64).eps * 4.0


def euler_to_mat(euler):
    """R = Rz(e[2]) @ Ry(e[1]) @ Rx(e[0])."""
    euler = np.asarray(euler, dtype=np.float64)
    a, b, c = euler[..., 0], euler[..., 1], euler[..., 2]
    sa, ca, sb, cb, sc, cc = np.sin(a), np.cos(a), np.sin(b), np.cos(b), np.sin(c), np.cos(c)
    mat = np.empty(euler.shape[:-1] + (3, 3), dtype=np.float64)
    mat[..., 0, 0] = cb * cc
    mat[..., 0, 1] = sa * sb * cc - ca * sc
    mat[..., 0, 2] = ca * sb * cc + sa * sc
    mat[..., 1, 0] = cb * sc
    mat[..., 1, 1] = sa * sb * sc + ca * cc
    mat[..., 1, 2] = ca * sb * sc - sa * cc
    mat[..., 2, 0] = -sb
    mat[..., 2, 1] = sa * cb
    mat[..., 2, 2] = ca * cb
    return mat


def mat_to_euler(mat):
    mat = np.asarray(mat, dtype=np.float64)
    cy = np.sqrt(mat[..., 0, 0] ** 2 + mat[..., 1, 0] ** 2)
    ok = cy > _EPS4
    euler = np.empty(mat.shape[:-1], dtype=np.float64)
    euler[..., 2] = np.where(ok, np.arctan2(mat[..., 1, 0], mat[..., 0, 0]), 0.0)
    euler[..., 1] = np.arctan2(-mat[..., 2, 0], cy)
    euler[..., 0] = np.where(ok, np.arctan2(mat[..., 2, 1], mat[..., 2, 2]), np.arctan2(-mat[..., 1, 2], mat[..., 1, 1]))
    return euler


def rotate_obs(state, angle, idx_rot, idx_xvel, idx_yvel):
    """Rotate a state (or batch of states) about the vertical axis (reference: utils/math.py:5-30)."""
    state = np.array(state)
    out = state.copy()
    out[idx_rot] = (state[idx_rot] + angle + np.pi) % (2 * np.pi) - np.pi
    out[idx_xvel] = np.cos(angle) * state[idx_xvel] - np.sin(angle) * state[idx_yvel]
    out[idx_yvel] = np.sin(angle) * state[idx_xvel] + np.cos(angle) * state[idx_yvel]
    return out


def mat2angle_xy(mat):
    return mat_to_euler(np.asarray(mat).reshape((3, 3)))[-1]


def angle2mat_xy(angle):
    return euler_to_mat(np.array([0, 0, angle]))


def transform_angle_2pi(angle):
    return (angle + np.pi) % (2 * np.pi) - np.pi
