/*
 * locosim C-ABI: the drop-in boundary of the B200 batched LocoEnv.step() engine.
 *
 * The reference has no FFI: its hot path is reached through the Python class hierarchy
 *   LocoEnv.step/reset                      /root/reference/loco_mujoco/environments/base.py:25,178-203 (step inherited
 *                                           from mushroom_rl MultiMuJoCo -> mujoco.mj_step, parameters base.py:32-33,109-111)
 *   GymnasiumWrapper.step/reset             /root/reference/loco_mujoco/environments/gymnasium.py:47-77
 * Each entry point below names the reference call it replaces. The Python facade
 * (loco_mujoco_b200/environments/base.py) binds these with ctypes; INTEGRATION.md shows the stub a maintainer of the
 * reference would add.
 *
 * Conventions: plain pointers + sizes, no C++/torch types. All `d_*` pointers are DEVICE pointers owned by the
 * caller (e.g. torch.cuda tensors), row-major, env-major ([n_envs, dim]). All work is enqueued on `stream`
 * (a cudaStream_t passed as void*; NULL = default stream); calls are CUDA-graph capturable; no allocation happens
 * in step/reset. Return value: 0 = OK, non-zero = error, message via locosim_last_error(). Never throws.
 * One handle per device; calls on one handle must be externally serialised.
 */
#ifndef LOCOSIM_H
#define LOCOSIM_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct locosim_handle locosim_t;

/* Replaces LocoEnv.__init__ -> MultiMuJoCo.__init__ (model compile + MjData) and load_trajectory
 * (base.py:31-143,145-168).  `model_*`: ModelPack blobs (include/locosim_modelpack.h); `task_*`: TaskSpec blobs
 * (include/locosim_task.h).  env_id_offset: global index of this shard's first env (multi-GPU sharding keeps the
 * per-env random streams identical to the single-GPU run). */
int locosim_create(const int32_t* model_ints, int n_model_ints, const double* model_reals, int n_model_reals,
                   const int32_t* task_ints, int n_task_ints, const double* task_reals, int n_task_reals,
                   int n_envs, int device, uint64_t seed, int64_t env_id_offset, locosim_t** out);
void locosim_destroy(locosim_t* h);
const char* locosim_last_error(const locosim_t* h);   /* h may be NULL: error of the last failed create */

int locosim_num_envs(const locosim_t* h);
int locosim_obs_dim(const locosim_t* h);
int locosim_action_dim(const locosim_t* h);
int locosim_nq(const locosim_t* h);

/* Newton-solver controls of the fp32 engine (defaults: tolerance 1e-5, ls_tolerance 0.1, max_iter 20, ls_iter 16) */
int locosim_set_solver(locosim_t* h, float tolerance, float ls_tolerance, int max_iter, int ls_iter);

/* Replaces LocoEnv.reset() (base.py:178-203 -> setup :205-241 -> Trajectory.reset_trajectory trajectory.py:236-273
 * -> set_sim_state :478-497).  d_mask: uint8 [n_envs] (NULL = all).  Draws (traj, sample) per env from the
 * counter-based stream unless d_traj_no / d_step_no (int32 [n_envs], may be NULL) pin them.  Writes the reset
 * observation to d_obs ([n_envs, obs_dim], may be NULL). */
int locosim_reset(locosim_t* h, const uint8_t* d_mask, const int32_t* d_traj_no, const int32_t* d_step_no, float* d_obs,
                  void* stream);
/* Same, with the parameter-pool row (= the model of a multi-model env, `np.random.randint(0, len(self._models))` in
 * base.py:187-191) pinned per env as well: d_pool_row int32 [n_envs] or NULL (draw from the counter-based stream). */
int locosim_reset_rows(locosim_t* h, const uint8_t* d_mask, const int32_t* d_traj_no, const int32_t* d_step_no,
                       const int32_t* d_pool_row, float* d_obs, void* stream);

/* Replaces LocoEnv.step(action) (mushroom MuJoCo.step: _preprocess_action base.py:606-621, n_substeps x mj_step,
 * _create_observation :584-604, is_absorbing :243-255, reward :170-176) for every env of the batch, followed by an
 * in-kernel auto-reset of the envs that terminated.
 *   d_action   [n_envs, action_dim] in [-1, 1]
 *   d_obs      [n_envs, obs_dim]   observation after the step (the terminal observation for envs with done=1)
 *   d_reward   [n_envs]
 *   d_done     [n_envs] uint8      absorbing flag (has_fallen, or non-finite state)
 *   d_next_obs [n_envs, obs_dim]   observation to act on next (== d_obs unless done, then the reset observation); may be NULL
 * auto_reset: 0 leaves terminated envs in their terminal state (caller resets with locosim_reset). */
int locosim_step(locosim_t* h, const float* d_action, float* d_obs, float* d_reward, uint8_t* d_done, float* d_next_obs,
                 int auto_reset, void* stream);

/* State access (LocoEnv.set_sim_state base.py:478-497 / data.qpos, data.qvel). fp32 [n_envs, nq]. */
int locosim_get_state(locosim_t* h, float* d_qpos, float* d_qvel, float* d_qacc_warmstart, void* stream);
int locosim_set_state(locosim_t* h, const float* d_qpos, const float* d_qvel, const float* d_qacc_warmstart, void* stream);

/* setup_random_rot in the drop-in single-env mode (unitreeA1.py:270-285: angle = np.random.uniform(0, 2 pi) from the legacy
 * numpy stream): rotation angles for the NEXT locosim_reset* call only, d_angle fp32 [n_envs] (NULL: the engine's own
 * counter-based draw, which is also what in-kernel auto-resets use). Requires a TaskSpec with TKI_ROT_* set. */
int locosim_set_reset_rotation(locosim_t* h, const float* d_angle);

/* Trajectory cursor of every env (traj * traj_len + sample; LS_REWARD_TRACKING, include/locosim_task.h). int32 [n_envs]. */
int locosim_get_cursor(locosim_t* h, int32_t* d_out, void* stream);

/* Replaces LocoEnv.create_dataset() (base.py:278-312 -> utils/trajectory.py:104-151) on the device: consecutive samples of
 * every trajectory of the reset table in observation layout. d_states / d_next_states fp32 [locosim_dataset_rows, obs_dim],
 * d_last fp32 [locosim_dataset_rows] (may be NULL); absorbing is all zero in the reference and is not materialised. */
int locosim_dataset_rows(const locosim_t* h);
int locosim_create_dataset(locosim_t* h, float* d_states, float* d_next_states, float* d_last, void* stream);

/* Per-episode goal features (UnitreeA1: cos / sin of the goal direction and the goal speed, GoalDirectionVelocity set in
 * setup() unitreeA1.py:287-291); normally loaded from the reset table, settable for reset(obs=...) (base.py:217-218,633-654).
 * d_goal fp32 [n_envs, 4]. */
int locosim_set_goal(locosim_t* h, const float* d_goal, void* stream);

/* Diagnostics: per-env counters since create: [0]=env steps, [1]=resets, [2]=Newton iterations of the last control
 * step (summed over its sub-steps; also the regrouping key), [3]=contacts of the last sub-step, [4]=terminations caused by
 * a non-finite state, [5..7]=max over control steps of [2] / the last sub-step's contacts / constraint rows.  d_out int32 [n_envs, 8]. */
int locosim_get_counters(locosim_t* h, int32_t* d_out, void* stream);

/* Domain randomisation (replaces DomainRandomizationHandler + per-reset MjModel recompilation,
 * /root/reference/loco_mujoco/utils/domain_randomization.py:191-296, hooked at base.py:183-185): a HOST pool of n_rows
 * parameter sets (row layout: loco_mujoco_b200/domain_randomization.py POOL_FIELDS + meaninertia + 4 user features, row_len
 * floats as float64) is uploaded once; the same mechanism carries multi-model envs (carry tasks: one row per weight,
 * user feature 0 = the weight's mass, observed through LS_OBS_PARAM); every (auto-)reset draws one row per env from the engine's counter-based stream. */
int locosim_param_pool_row_len(const locosim_t* h);
int locosim_set_param_pool(locosim_t* h, const double* pool, int n_rows, int row_len);
/* pool row currently used by each env: d_out int32 [n_envs] */
int locosim_get_param_rows(locosim_t* h, int32_t* d_out, void* stream);

/* Kernels enqueued by one locosim_step: the fused step kernel, preceded (default) by the one-block regrouping kernel that
 * buckets the envs by the solver effort of their previous step (scheduling only; results do not depend on it). */
int locosim_kernels_per_step(const locosim_t* h);

/* Measurement aid (bench.py): non-tensor FP32 FMA throughput of `device` in TFLOP/s, measured with a register-only
 * FMA-chain kernel on all SMs (best of 5). Not part of the simulation path. */
int locosim_measure_fp32_peak(int device, double* tflops_out);

/* Diagnostics of the convex (MPR) narrow phase, process-wide, only counted when the library was created with the environment
 * variable LOCOSIM_DEBUG having bit 8 set: out8[0] = MPR runs, [1] = support calls, [3] = runs that hit the iteration cap. */
int locosim_debug_counters(unsigned long long* out8);

/* Launch geometry chosen for this handle: warps(envs) per block, dynamic shared memory bytes per block, blocks. */
int locosim_launch_info(const locosim_t* h, int* warps_per_block, int* smem_bytes, int* n_blocks);

#ifdef __cplusplus
}
#endif
#endif
