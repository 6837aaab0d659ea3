"""
ctypes binding of the CUDA engine's C-ABI (include/locosim.h -> loco_mujoco_b200/liblocosim_cuda.so).

There is NO CPU fallback: if the shared library is missing or no CUDA device is visible, constructing an engine
raises. (The fp64 CPU oracle under oracle/ is test infrastructure and is never imported from here.)
"""
import ctypes
import os

import numpy as np

_LIB = None
_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "liblocosim_cuda.so")


class EngineUnavailable(RuntimeError):
    pass


def load_library():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(_LIB_PATH):
        raise EngineUnavailable("CUDA engine library not built: %s (run `python -c 'import __graft_entry__ as g; "
                                "g.build()'`)" % _LIB_PATH)
    lib = ctypes.CDLL(_LIB_PATH)
    vp, ip = ctypes.c_void_p, ctypes.c_int
    lib.locosim_create.restype = ip
    lib.locosim_create.argtypes = [vp, ip, vp, ip, vp, ip, vp, ip, ip, ip, ctypes.c_uint64, ctypes.c_int64,
                                   ctypes.POINTER(vp)]
    lib.locosim_destroy.restype = None
    lib.locosim_destroy.argtypes = [vp]
    lib.locosim_last_error.restype = ctypes.c_char_p
    lib.locosim_last_error.argtypes = [vp]
    for name in ("locosim_num_envs", "locosim_obs_dim", "locosim_action_dim", "locosim_nq"):
        getattr(lib, name).restype = ip
        getattr(lib, name).argtypes = [vp]
    lib.locosim_set_solver.restype = ip
    lib.locosim_set_solver.argtypes = [vp, ctypes.c_float, ctypes.c_float, ip, ip]
    lib.locosim_reset.restype = ip
    lib.locosim_reset.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.locosim_reset_rows.restype = ip
    lib.locosim_reset_rows.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.locosim_step.restype = ip
    lib.locosim_step.argtypes = [vp, vp, vp, vp, vp, vp, ip, vp]
    lib.locosim_get_state.restype = ip
    lib.locosim_get_state.argtypes = [vp, vp, vp, vp, vp]
    lib.locosim_set_state.restype = ip
    lib.locosim_set_state.argtypes = [vp, vp, vp, vp, vp]
    lib.locosim_set_reset_rotation.restype = ip
    lib.locosim_set_reset_rotation.argtypes = [vp, vp]
    lib.locosim_get_cursor.restype = ip
    lib.locosim_get_cursor.argtypes = [vp, vp, vp]
    lib.locosim_dataset_rows.restype = ip
    lib.locosim_dataset_rows.argtypes = [vp]
    lib.locosim_create_dataset.restype = ip
    lib.locosim_create_dataset.argtypes = [vp, vp, vp, vp, vp]
    lib.locosim_set_goal.restype = ip
    lib.locosim_set_goal.argtypes = [vp, vp, vp]
    lib.locosim_get_counters.restype = ip
    lib.locosim_get_counters.argtypes = [vp, vp, vp]
    lib.locosim_param_pool_row_len.restype = ip
    lib.locosim_param_pool_row_len.argtypes = [vp]
    lib.locosim_set_param_pool.restype = ip
    lib.locosim_set_param_pool.argtypes = [vp, vp, ip, ip]
    lib.locosim_get_param_rows.restype = ip
    lib.locosim_get_param_rows.argtypes = [vp, vp, vp]
    lib.locosim_kernels_per_step.restype = ip
    lib.locosim_kernels_per_step.argtypes = [vp]
    lib.locosim_measure_fp32_peak.restype = ip
    lib.locosim_measure_fp32_peak.argtypes = [ip, ctypes.POINTER(ctypes.c_double)]
    lib.locosim_launch_info.restype = ip
    lib.locosim_launch_info.argtypes = [vp, ctypes.POINTER(ip), ctypes.POINTER(ip), ctypes.POINTER(ip)]
    _LIB = lib
    return lib


EXPORTED_SYMBOLS = ["locosim_create", "locosim_destroy", "locosim_last_error", "locosim_num_envs", "locosim_obs_dim",
                    "locosim_action_dim", "locosim_nq", "locosim_set_solver", "locosim_reset", "locosim_step",
                    "locosim_get_state", "locosim_set_state", "locosim_get_counters", "locosim_launch_info",
                    "locosim_param_pool_row_len", "locosim_set_param_pool", "locosim_get_param_rows",
                    "locosim_kernels_per_step", "locosim_reset_rows", "locosim_set_goal", "locosim_measure_fp32_peak",
                    "locosim_set_reset_rotation", "locosim_get_cursor", "locosim_dataset_rows", "locosim_create_dataset",
                    "locosim_debug_counters"]


def measure_fp32_peak(device=0):
    """Measured non-tensor FP32 FMA peak of the device in TFLOP/s (include/locosim.h)."""
    v = ctypes.c_double(0)
    if load_library().locosim_measure_fp32_peak(int(device), ctypes.byref(v)) != 0:
        raise RuntimeError("locosim_measure_fp32_peak failed")
    return v.value


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class CudaEngine:
    """n_envs independent worlds of one compiled model on one GPU; all I/O are torch.cuda tensors."""

    def __init__(self, model_blobs, task_blobs, n_envs, device=0, seed=0, env_id_offset=0, warps_per_block=None):
        import torch
        if not torch.cuda.is_available():
            raise EngineUnavailable("no CUDA device visible: the locosim engine has no CPU path")
        self.torch = torch
        self.lib = load_library()
        mi, mr = model_blobs
        ti, tr = task_blobs
        mi = np.ascontiguousarray(mi, dtype=np.int32)
        mr = np.ascontiguousarray(mr, dtype=np.float64)
        ti = np.ascontiguousarray(ti, dtype=np.int32)
        tr = np.ascontiguousarray(tr, dtype=np.float64)
        h = ctypes.c_void_p()
        # launch geometry: envs (= warps) per block; None = the engine's own choice (the largest block that keeps the most
        # envs resident per SM). Smaller blocks let several engines share an SM (MixedBatch). Scheduling only.
        prev = os.environ.get("LOCOSIM_WPB")
        if warps_per_block is not None:
            os.environ["LOCOSIM_WPB"] = str(int(warps_per_block))
        try:
            rc = self._create(mi, mr, ti, tr, n_envs, device, seed, env_id_offset, h)
        finally:
            if warps_per_block is not None:
                if prev is None:
                    os.environ.pop("LOCOSIM_WPB", None)
                else:
                    os.environ["LOCOSIM_WPB"] = prev
        if rc != 0:
            raise RuntimeError("locosim_create failed: %s" % self.lib.locosim_last_error(None).decode())
        self._finish_init(h, n_envs, device)

    def _create(self, mi, mr, ti, tr, n_envs, device, seed, env_id_offset, h):
        return self.lib.locosim_create(mi.ctypes.data, len(mi), mr.ctypes.data, len(mr), ti.ctypes.data, len(ti),
                                       tr.ctypes.data, len(tr), int(n_envs), int(device), int(seed), int(env_id_offset),
                                       ctypes.byref(h))

    def _finish_init(self, h, n_envs, device):
        torch = self.torch
        self.h = h
        self.device = torch.device("cuda", device)
        self.n_envs = int(n_envs)
        self.obs_dim = self.lib.locosim_obs_dim(h)
        self.action_dim = self.lib.locosim_action_dim(h)
        self.nq = self.lib.locosim_nq(h)
        # obs / reward / done of a step live in ONE device allocation (three views), so that a consumer on the host can
        # fetch the whole step result with a single D2H copy of `packed_out`
        N, D = int(n_envs), self.obs_dim
        self.packed_out = torch.zeros((N * (4 * D + 5),), dtype=torch.uint8, device=self.device)
        self.obs, self.reward, self.done = self.views_of(self.packed_out)
        self.next_obs = torch.zeros((n_envs, self.obs_dim), dtype=torch.float32, device=self.device)
        self.launches = 0
        self.kernels_per_step = self.lib.locosim_kernels_per_step(h)

    def views_of(self, packed):
        """(obs [N, D] f32, reward [N] f32, done [N] u8) views of a buffer laid out like `packed_out`."""
        N, D, t = self.n_envs, self.obs_dim, self.torch
        return (packed[:4 * N * D].view(t.float32).view(N, D), packed[4 * N * D:4 * N * D + 4 * N].view(t.float32),
                packed[4 * N * D + 4 * N:])

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("locosim: %s" % self.lib.locosim_last_error(self.h).decode())

    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, "h", None):
            self.lib.locosim_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_solver(self, tolerance=1e-5, ls_tolerance=0.1, max_iter=20, ls_iter=16):
        self._check(self.lib.locosim_set_solver(self.h, tolerance, ls_tolerance, max_iter, ls_iter))

    def reset(self, mask=None, traj_no=None, step_no=None, out=None, pool_row=None, rot_angle=None):
        """rot_angle: float32 cuda [n_envs], rotation angles of this reset (setup_random_rot with host-drawn angles)."""
        out = self.next_obs if out is None else out
        if rot_angle is not None:
            self._check(self.lib.locosim_set_reset_rotation(self.h, _ptr(rot_angle)))
        self._check(self.lib.locosim_reset_rows(self.h, _ptr(mask), _ptr(traj_no), _ptr(step_no), _ptr(pool_row), _ptr(out),
                                                self._stream()))
        self.launches += 1
        return out

    @property
    def packed_bytes(self):
        """Bytes of one step's (obs f32 [N, D] | reward f32 [N] | done u8 [N]) record."""
        return self.n_envs * (4 * self.obs_dim + 5)

    def step(self, action, auto_reset=True, want_next_obs=True, packed=None):
        """action: float32 cuda [n_envs, action_dim] (contiguous). Returns (obs, reward, done, next_obs) views.
        packed: optional uint8 cuda buffer of `packed_bytes` (16-byte aligned, e.g. one time slot of a rollout buffer)
        the kernel writes this step's obs / reward / done into instead of the engine's own `packed_out`."""
        if action.dtype != self.torch.float32 or not action.is_contiguous() or action.device != self.device:
            raise ValueError("action must be a contiguous float32 tensor on %s" % self.device)
        if tuple(action.shape) != (self.n_envs, self.action_dim):
            raise ValueError("action shape %s != %s" % (tuple(action.shape), (self.n_envs, self.action_dim)))
        if packed is None:
            obs, reward, done = self.obs, self.reward, self.done
        else:
            if packed.dtype != self.torch.uint8 or packed.numel() != self.packed_bytes or not packed.is_contiguous() \
                    or packed.data_ptr() % 4:
                raise ValueError("packed must be a contiguous, 4-byte aligned uint8 buffer of %d bytes" % self.packed_bytes)
            obs, reward, done = self.views_of(packed)
        self._check(self.lib.locosim_step(self.h, _ptr(action), _ptr(obs), _ptr(reward), _ptr(done),
                                          _ptr(self.next_obs) if want_next_obs else None, int(bool(auto_reset)),
                                          self._stream()))
        self.launches += self.kernels_per_step
        return obs, reward, done, self.next_obs

    def get_state(self):
        t = self.torch
        q = t.empty((self.n_envs, self.nq), dtype=t.float32, device=self.device)
        v = t.empty_like(q)
        w = t.empty_like(q)
        self._check(self.lib.locosim_get_state(self.h, _ptr(q), _ptr(v), _ptr(w), self._stream()))
        return q, v, w

    def set_state(self, qpos, qvel, qacc_warmstart=None):
        for x in (qpos, qvel, qacc_warmstart):
            if x is not None and (x.dtype != self.torch.float32 or not x.is_contiguous()):
                raise ValueError("state tensors must be contiguous float32")
        self._check(self.lib.locosim_set_state(self.h, _ptr(qpos), _ptr(qvel), _ptr(qacc_warmstart), self._stream()))

    def cursor(self):
        t = self.torch
        c = t.empty((self.n_envs,), dtype=t.int32, device=self.device)
        self._check(self.lib.locosim_get_cursor(self.h, _ptr(c), self._stream()))
        return c

    def create_dataset(self):
        """(states, next_states, last) of the reset table in observation layout, built on the device."""
        t = self.torch
        n = self.lib.locosim_dataset_rows(self.h)
        states = t.empty((n, self.obs_dim), dtype=t.float32, device=self.device)
        nxt = t.empty_like(states)
        last = t.empty((n,), dtype=t.float32, device=self.device)
        self._check(self.lib.locosim_create_dataset(self.h, _ptr(states), _ptr(nxt), _ptr(last), self._stream()))
        self.launches += 1
        return states, nxt, last

    def set_goal(self, goal):
        """goal: float32 cuda [n_envs, 4] per-episode goal features (A1: cos, sin, speed)."""
        if goal.dtype != self.torch.float32 or not goal.is_contiguous() or tuple(goal.shape) != (self.n_envs, 4):
            raise ValueError("goal must be a contiguous float32 [n_envs, 4] tensor")
        self._check(self.lib.locosim_set_goal(self.h, _ptr(goal), self._stream()))

    def set_param_pool(self, pool):
        """pool: float64 [n_rows, row_len] (domain_randomization.pool_row layout)."""
        pool = np.ascontiguousarray(pool, dtype=np.float64)
        if pool.ndim != 2 or pool.shape[1] != self.lib.locosim_param_pool_row_len(self.h):
            raise ValueError("pool row length %s != %d" % (pool.shape, self.lib.locosim_param_pool_row_len(self.h)))
        self._check(self.lib.locosim_set_param_pool(self.h, pool.ctypes.data, pool.shape[0], pool.shape[1]))

    def param_rows(self):
        t = self.torch
        r = t.empty((self.n_envs,), dtype=t.int32, device=self.device)
        self._check(self.lib.locosim_get_param_rows(self.h, _ptr(r), self._stream()))
        return r

    def counters(self):
        t = self.torch
        c = t.empty((self.n_envs, 8), dtype=t.int32, device=self.device)
        self._check(self.lib.locosim_get_counters(self.h, _ptr(c), self._stream()))
        return c

    def launch_info(self):
        a, b, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self.lib.locosim_launch_info(self.h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
        return dict(warps_per_block=a.value, smem_bytes=b.value, blocks=c.value)
