// locosim.cu -- CUDA engine + C-ABI (include/locosim.h) of the batched LocoEnv.step() hot path, sm_100a.
//
// One fused kernel per control step: load state (HBM, env-major, coalesced per warp) -> n_substeps x
// [kinematics, CRB, collision, constraint assembly, Newton solve, integrate] entirely in shared memory ->
// observation gather + has_fallen + reward -> in-kernel auto-reset from the trajectory table -> store.
// See locosim_core.cuh for the per-warp algorithm and DESIGN.md for the layout / roofline accounting.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/locosim.h"
#include "locosim_config.h"
#include "locosim_host.h"

// ----------------------------------------------------------------------------------------------------------
struct EngineState {
  float *qpos, *qvel, *ws, *goal;   // [N,nv] [N,nv] [N,nv] [N,4]
  int* episode;                     // [N] reset counter (drives the per-env random stream)
  int* counters;                    // [N,8]
  int* dr_row;                      // [N] row of the parameter pool used by the env's current episode
  int* perm;                        // [N] env handled by warp slot i (regrouped every step by solver effort)
  int* cursor;                      // [N] trajectory cursor traj * traj_len + sample (LS_REWARD_TRACKING)
  const float* rot_angle;           // [N] rotation angles pinned for the next reset (setup_random_rot, drop-in mode) or NULL
  const float* pool;                // [K, P] parameter pool (domain randomisation); K = 1: the model's own values
  int pool_K;
};

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
// counter-based draw of (trajectory, sample) for reset number `episode` of global env `genv`
__device__ __forceinline__ void draw_reset(uint64_t seed, int64_t genv, int episode, int n_traj, int traj_len, int* traj,
                                           int* step) {
  uint64_t r = mix64(seed ^ mix64((uint64_t)genv * 0x100000001B3ULL + (uint64_t)(uint32_t)episode));
  *traj = (int)((r & 0xffffffffULL) % (uint64_t)n_traj);
  *step = (int)((r >> 32) % (uint64_t)traj_len);
}

// setup_random_rot: rotation angle ~ U[0, 2 pi) of reset number `episode` of global env `genv`
__device__ __forceinline__ float draw_rot_angle(uint64_t seed, int64_t genv, int episode) {
  uint64_t r = mix64((seed + 0x2545F4914F6CDD1DULL) ^ mix64((uint64_t)genv * 0x100000001B3ULL + (uint64_t)(uint32_t)episode));
  return (float)(r >> 40) * (6.283185307179586f / 16777216.0f);
}

__device__ __forceinline__ int draw_pool_row(uint64_t seed, int64_t genv, int episode, int K) {
  if (K <= 1) return 0;
  uint64_t r = mix64((seed + 0x5DEECE66DULL) ^ mix64((uint64_t)genv * 0x100000001B3ULL + (uint64_t)(uint32_t)episode));
  return (int)(r % (uint64_t)K);
}

// ----------------------------------------------------------------------------------------------------------
// reset kernel: one warp per env, no shared memory
// ----------------------------------------------------------------------------------------------------------
__global__ void reset_kernel(int ms, DevTask t, EngineState st, const uint8_t* mask, const int* traj_no,
                             const int* step_no, const int* pool_row, float* obs, int n_envs, uint64_t seed,
                             int64_t env_off) {
  int env = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (env >= n_envs) return;
  if (mask && !mask[env]) return;
  const DevModel& m = c_models[ms];
  const int nv = m.nv, ncol = 2 * nv + t.n_goal;
  int tr, sp;
  int ep = st.episode[env];
  draw_reset(seed, env_off + env, ep, t.n_traj, t.traj_len, &tr, &sp);
  if (traj_no) tr = min(max(traj_no[env], 0), t.n_traj - 1);       // (pinned rows are clamped into the table)
  if (step_no) sp = min(max(step_no[env], 0), t.traj_len - 1);
  const float* row = t.table + ((size_t)tr * t.traj_len + sp) * ncol;
  int prow = draw_pool_row(seed, env_off + env, ep, st.pool_K);     // the model of this episode (base.py:187-191)
  if (pool_row) prow = min(max(pool_row[env], 0), st.pool_K - 1);
  float angle = 0.0f;
  if (t.rot[0] >= 0) angle = st.rot_angle ? st.rot_angle[env] : draw_rot_angle(seed, env_off + env, ep);
  for (int i = lane; i < nv; i += 32) {
    st.qpos[(size_t)env * nv + i] = reset_value(t, row, nv, LS_OBS_QPOS, i, angle);
    st.qvel[(size_t)env * nv + i] = reset_value(t, row, nv, LS_OBS_QVEL, i, angle);
    st.ws[(size_t)env * nv + i] = 0;
  }
  for (int k = lane; k < t.n_goal; k += 32) st.goal[(size_t)env * 4 + k] = row[2 * nv + k];
  if (obs) {
    for (int k = lane; k < t.obs_dim; k += 32) {
      int idx = t.obs_src_idx[k], ty = t.obs_src_type[k];
      float v;
      if (ty == LS_OBS_QPOS || ty == LS_OBS_QVEL) v = reset_value(t, row, nv, ty, idx, angle);
      else if (ty == LS_OBS_GRF) v = 0.0f;           // the running mean of the foot forces is reset with the episode
      else if (ty == LS_OBS_PARAM) v = st.pool[(size_t)prow * m.pool_P + t.po_user + idx];
      else v = row[2 * nv + idx];
      obs[(size_t)env * t.obs_dim + k] = v;
    }
  }
  if (lane == 0) {
    st.episode[env] = ep + 1; st.counters[(size_t)env * 8 + 1] += 1;
    st.dr_row[env] = prow;
    st.cursor[env] = tr * t.traj_len + sp;
  }
}

// ----------------------------------------------------------------------------------------------------------
// step kernel: one warp per env, EnvS in dynamic shared memory
// ----------------------------------------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(512, 1) step_kernel(int ms, DevTask t, SolverOpts so, EngineState st,
                                                    const float* __restrict__ action, float* __restrict__ obs,
                                                    float* __restrict__ reward, uint8_t* __restrict__ done,
                                                    float* __restrict__ next_obs, int n_envs, int auto_reset,
                                                    uint64_t seed, int64_t env_off, int sync_substeps, int key_mode_in,
                                                    const unsigned char* __restrict__ stage_src, int stage_bytes) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int key_mode = key_mode_in & 0xffff;
  if (stage_bytes > 0) {
    // LOCOSIM_STAGE (default on where it fits): the table of the PRIMITIVE candidate pairs of the mid-phase (packed geom pair +
    // bound; UnitreeA1: 357 entries read in every dynamics evaluation by every warp) is staged once per block into the shared
    // memory behind the per-env working sets by ONE TMA bulk copy (cp.async.bulk global -> shared::cta, completion on an
    // mbarrier); EnvS::pk_tab points to it, DevModel::pk_n is its entry count.
    unsigned char* dst = smem_raw + (size_t)(blockDim.x >> 5) * sizeof(EnvS<C>);
    const unsigned dst_a = (unsigned)__cvta_generic_to_shared(dst);
    const unsigned bar_a = dst_a + (unsigned)stage_bytes;
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a) : "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"((unsigned)stage_bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_a),
                   "l"(stage_src), "r"((unsigned)stage_bytes), "r"(bar_a)
                   : "memory");
    }
    __syncthreads();                                    // (the barrier object is initialised before anybody polls it)
    unsigned ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(bar_a) : "memory");
  }
  // warps past the last env shadow it (same barriers, no stores): the block-wide barriers inside the solver have
  // data-dependent counts, so every warp of a block has to run the physics
  const int slot_raw = blockIdx.x * (blockDim.x >> 5) + warp;
  const bool ghost = slot_raw >= n_envs;
  const int env = st.perm[ghost ? n_envs - 1 : slot_raw];
  EnvS<C>& e = reinterpret_cast<EnvS<C>*>(smem_raw)[warp];
  const DevModel& m = c_models[ms];
  const int nv = m.nv, nu = m.nu, D = t.obs_dim;
  if (lane == 0) {
    e.prm = st.pool + (size_t)st.dr_row[env] * m.pool_P;
    if (stage_bytes > 0) {
      const unsigned char* sm = smem_raw + (size_t)(blockDim.x >> 5) * sizeof(EnvS<C>);
      e.pk_tab = reinterpret_cast<const int*>(sm);
    } else { e.pk_tab = m.pair_packed; }
  }

  // ---- load state ----
  for (int i = lane; i < nv; i += 32) {
    e.qpos[i] = st.qpos[(size_t)env * nv + i];
    e.qvel[i] = st.qvel[(size_t)env * nv + i];
    e.qacc_ws[i] = st.ws[(size_t)env * nv + i];
    e.qacc[i] = 0;
  }
  for (int i = lane; i < nu; i += 32) e.ctrl[t.act_idx[i]] = action[(size_t)env * nu + i] * t.act_delta[i] + t.act_mean[i];
  if (lane < 4) e.goal[lane] = st.goal[(size_t)env * 4 + lane];
  __syncwarp();
  init_workspace(ms, e);
  if (lane == 0) e.iter_sum = 0;
  if (lane < 3 * LS_MAX_GRF) e.grf[lane] = 0;
  // reward depends on the *previous* observation only (utils/reward.py via base.py:170-176): evaluate it now
  float rew = 0;
  if (t.reward_type == LS_REWARD_TARGET_VELOCITY) { float d = obs_value(t, e, t.ri[0]) - t.rp[0]; rew = expf(-d * d); }
  else if (t.reward_type == LS_REWARD_VELOCITY_VECTOR) {
    float g = obs_value(t, e, t.ri[3]);
    float dx = obs_value(t, e, t.ri[0]) - g * obs_value(t, e, t.ri[2]);
    float dy = obs_value(t, e, t.ri[1]) - g * obs_value(t, e, t.ri[2] + 1);
    rew = expf(-5.0f * sqrtf(dx * dx + dy * dy));
  } else if (t.reward_type == LS_REWARD_POS) rew = obs_value(t, e, t.ri[0]);

  // ---- physics ----
  if (sync_substeps) {
    // keep the warps of a block in the same phase: they then share instruction-cache lines
    for (int k = 0; k < t.n_substeps; k++) { if (k % sync_substeps == 0) __syncthreads(); physics_substeps(ms, e, so, 1, t.grf_group, t.n_grf); }
  } else {
    physics_substeps(ms, e, so, t.n_substeps, t.grf_group, t.n_grf);
  }

  if (ghost) return;

  // ---- observation, termination, reward ----
  bool bad = false;
  for (int i = lane; i < nv; i += 32) bad |= !(isfinite(e.qpos[i]) && isfinite(e.qvel[i]));
  bool fallen = false;
  for (int k = lane; k < t.n_done; k += 32) {
    float v = obs_value(t, e, t.done_obs_idx[k]);
    fallen |= (v < t.done_lo[k]) || (v > t.done_hi[k]);
  }
  bad = __any_sync(0xffffffffu, bad);
  fallen = __any_sync(0xffffffffu, fallen) && t.use_absorbing;
  const bool is_done = fallen || bad;
  if (t.reward_type == LS_REWARD_TRACKING) {
    // mocap-tracking reward (include/locosim_task.h): the cursor advances one sample, the observation reached by this step
    // is compared with the table row at the advanced cursor
    int cur = st.cursor[env];
    const int tj = cur / t.traj_len;
    int sp = cur - tj * t.traj_len;
    sp = sp + 1 < t.traj_len ? sp + 1 : t.traj_len - 1;
    cur = tj * t.traj_len + sp;
    const float* ref = t.table + (size_t)cur * (2 * nv + t.n_goal);
    float ep = 0, ev = 0;
    for (int k = lane; k < D; k += 32) {
      const int ty = t.obs_src_type[k], idx = t.obs_src_idx[k];
      if (ty == LS_OBS_QPOS) { const float d = e.qpos[idx] - ref[idx]; ep += d * d; }
      else if (ty == LS_OBS_QVEL) { const float d = e.qvel[idx] - ref[nv + idx]; ev += d * d; }
    }
    ep = warp_sum(ep); ev = warp_sum(ev);
    rew = bad ? 0.0f : t.track[0] * expf(-t.track[1] * ep) + t.track[2] * expf(-t.track[3] * ev);
    if (lane == 0) st.cursor[env] = cur;
  }
  for (int k = lane; k < D; k += 32) {
    float v = obs_value(t, e, k);
    if (bad) v = 0.0f;
    obs[(size_t)env * D + k] = v;
  }
  if (lane == 0) {
    float r = rew;
    reward[env] = r;
    done[env] = is_done ? 1 : 0;
    int* cnt = st.counters + (size_t)env * 8;
    cnt[0] += 1; cnt[2] = e.iter_sum; cnt[3] = e.ncon;
    if (bad) cnt[4] += 1;
    {
      int key = e.iter_sum;
      const int mpr_w = (key_mode >> 8) & 255;           // weight of one MPR run in Newton-iteration units
      key_mode &= 255;
      if (key_mode == 1) key = (cnt[5] + e.iter_sum + 1) >> 1;
      else if (key_mode == 2) key = e.solver_iter * 8;
      else if (key_mode == 3) key = e.iter_sum + (e.nefc >> 1) + mpr_w * min(e.mpr_calls, 64);
      else if (key_mode == 4) key = (3 * cnt[5] + e.iter_sum + 2) >> 2;
      else if (key_mode == 5) key = e.iter_sum + 4 * e.solver_iter;
      else if (key_mode == 6) key = e.iter_sum + e.nefc;
      else if (key_mode == 7) key = e.iter_sum + 4 * e.solver_iter + (e.nefc >> 1);
      cnt[5] = key;
    } cnt[6] = max(cnt[6], e.ncon); cnt[7] = max(cnt[7], e.nefc);
  }

  // ---- auto-reset ----
  if (is_done && (auto_reset || bad)) {
    const int key_reset = (key_mode_in >> 16) ? (key_mode_in >> 16) - 1 : -1;
    int ep = st.episode[env];
    int tr, sp;
    draw_reset(seed, env_off + env, ep, t.n_traj, t.traj_len, &tr, &sp);
    __syncwarp();
    reset_env(ms, t, e, tr, sp, t.rot[0] >= 0 ? draw_rot_angle(seed, env_off + env, ep) : 0.0f);
    if (lane == 0) {
      const int prow = draw_pool_row(seed, env_off + env, ep, st.pool_K);
      st.episode[env] = ep + 1; st.counters[(size_t)env * 8 + 1] += 1;
      if (key_reset >= 0) st.counters[(size_t)env * 8 + 5] = key_reset;   // regrouping key of a fresh episode (the finished one's says nothing)
      st.dr_row[env] = prow;
      st.cursor[env] = tr * t.traj_len + sp;
      e.prm = st.pool + (size_t)prow * m.pool_P;      // the next episode's model (LS_OBS_PARAM entries of next_obs)
    }
    __syncwarp();
    for (int k = lane; k < t.n_goal; k += 32) st.goal[(size_t)env * 4 + k] = e.goal[k];
  }
  if (next_obs) for (int k = lane; k < D; k += 32) next_obs[(size_t)env * D + k] = obs_value(t, e, k);

  // ---- store state ----
  for (int i = lane; i < nv; i += 32) {
    st.qpos[(size_t)env * nv + i] = e.qpos[i];
    st.qvel[(size_t)env * nv + i] = e.qvel[i];
    st.ws[(size_t)env * nv + i] = e.qacc_ws[i];
  }
}

// ----------------------------------------------------------------------------------------------------------
// Regrouping: the warps of a block run the solver in lock-step, so a block is as slow as its hardest env. Solver effort
// is strongly correlated from one control step to the next (lag-1 correlation of the per-step Newton iteration count
// ~0.6), so before every step the envs are bucketed by the iteration count of their previous step (counting sort,
// one block) and envs of similar effort share a block. Pure scheduling: results do not depend on the mapping.
// ----------------------------------------------------------------------------------------------------------
#define LS_RANK_BUCKETS 128
__global__ void __launch_bounds__(1024) rank_kernel(const int* __restrict__ counters, int* __restrict__ perm, int n_envs,
                                                    int key_shift) {
  __shared__ int hist[LS_RANK_BUCKETS];
  __shared__ int base[LS_RANK_BUCKETS];
  for (int i = threadIdx.x; i < LS_RANK_BUCKETS; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n_envs; i += blockDim.x) {
    int k = counters[(size_t)i * 8 + 5] >> key_shift;
    atomicAdd(&hist[k < LS_RANK_BUCKETS - 1 ? k : LS_RANK_BUCKETS - 1], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int i = 0; i < LS_RANK_BUCKETS; i++) { base[i] = acc; acc += hist[i]; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_envs; i += blockDim.x) {
    int k = counters[(size_t)i * 8 + 5] >> key_shift;
    int pos = atomicAdd(&base[k < LS_RANK_BUCKETS - 1 ? k : LS_RANK_BUCKETS - 1], 1);
    perm[n_envs - 1 - pos] = i;     // hardest first: the long blocks start early, the short ones fill the tail
  }
}

// ----------------------------------------------------------------------------------------------------------
// handle
// ----------------------------------------------------------------------------------------------------------
struct locosim_handle {
  int device = 0, n_envs = 0, cfg = -1, wpb = 4, smem = 0, slot = -1, sync_substeps = 0, regroup = 1, key_shift = 0, key_mode = 3;
  uint64_t seed = 0;
  int64_t env_off = 0;
  HostModel hm;
  HostTask ht;
  DevModel dm;
  DevTask dt;
  SolverOpts so;
  EngineState st;
  int* d_mints = nullptr; float* d_mreals = nullptr; int* d_tints = nullptr; float* d_treals = nullptr;
  float* d_pool = nullptr;
  unsigned char* d_stage = nullptr; int stage_bytes = 0;
  std::string err;
};
static std::string g_create_error;
static bool g_slot_used[LS_MAX_SLOTS] = {false};

#define CK(call)                                                                      \
  do {                                                                                \
    cudaError_t _e = (call);                                                          \
    if (_e != cudaSuccess) {                                                          \
      h->err = std::string(#call) + ": " + cudaGetErrorString(_e);                    \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

static bool has_convex_pairs(const HostModel& hm) {
  const int* ip = hm.ints.data();
  // (pair_geom / geom_type live in the int blob in LOCOSIM_MP_INT_FIELDS order; resolve through a bound view)
  DevModel v;
  bind_model(v, hm, hm.ints.data(), hm.reals.data());
  (void)ip;
  for (int p = 0; p < hm.np; p++)
    if (v.pair_packed[p] & (2 << 24)) return true;
  return false;
}
static bool has_boxbox_pairs(const HostModel& hm) {
  DevModel v;
  bind_model(v, hm, hm.ints.data(), hm.reals.data());
  for (int p = 0; p < hm.np; p++)
    if (v.geom_type[v.pair_geom[2 * p]] == LS_GEOM_BOX && v.geom_type[v.pair_geom[2 * p + 1]] == LS_GEOM_BOX) return true;
  return false;
}
template <class C>
static bool cfg_fits(const HostModel& hm, const HostTask& ht) {
  if (!C::CONVEX && has_convex_pairs(hm)) return false;
  if (!C::BOXBOX && has_boxbox_pairs(hm)) return false;
  return hm.nv <= C::NV && hm.nb <= C::NB && hm.ng <= C::NG && ht.obs_dim <= C::MAXOBS && hm.cone == (int)C::CONE &&
         hm.integrator == (int)C::RK4;
}
template <class C>
static int cfg_smem() { return (int)sizeof(EnvS<C>); }
// sizeof(EnvS) decides the envs resident per SM (227 KB of dynamic shared memory per block on sm_100): a field added without
// looking costs a whole env per SM (and, for the RK4 robots, a third wave of blocks at 4096 envs). Guard the measured layout.
static_assert(sizeof(EnvS<CfgEllEuler>) * 15 <= 232448, "UnitreeA1: 15 envs per SM");
static_assert(sizeof(EnvS<CfgPyrEuler>) * 15 <= 232448, "Talos / UnitreeH1: 15 envs per SM");
static_assert(sizeof(EnvS<CfgPyrRK4>) * 14 <= 232448, "HumanoidTorque / Atlas: 14 envs per SM (293 blocks = 1.98 waves at 4096 envs)");
static_assert(sizeof(EnvS<CfgPyrEuler29>) * 9 <= 232448, "UnitreeG1: 9 envs per SM");

template <class C>
static int launch_step(locosim_handle* h, const float* a, float* o, float* r, uint8_t* d, float* no, int auto_reset,
                       cudaStream_t s) {
  int blocks = (h->n_envs + h->wpb - 1) / h->wpb;
  if (h->regroup) rank_kernel<<<1, 1024, 0, s>>>(h->st.counters, h->st.perm, h->n_envs, h->key_shift);
  step_kernel<C><<<blocks, h->wpb * 32, h->smem + (h->stage_bytes ? h->stage_bytes + 16 : 0), s>>>(
      h->slot, h->dt, h->so, h->st, a, o, r, d, no, h->n_envs, auto_reset, h->seed, h->env_off, h->sync_substeps, h->key_mode,
      h->d_stage, h->stage_bytes);
  CK(cudaGetLastError());
  return 0;
}
template <class C>
static int setup_cfg(locosim_handle* h) {
  int per_env = cfg_smem<C>();
  int dev_max = 0;
  CK(cudaDeviceGetAttribute(&dev_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
  int sm_total = 0;
  CK(cudaDeviceGetAttribute(&sm_total, cudaDevAttrMaxSharedMemoryPerMultiprocessor, h->device));
  // envs resident per SM for w warps per block (1 KB per block is reserved by the driver)
  auto envs_per_sm = [&](int w) {
    int sm = w * per_env;
    if (sm > dev_max) return 0;
    int b = sm_total / (sm + 1024);
    return (b > 32 ? 32 : b) * w;
  };
  int best = 1, best_env = 0;
  for (int w = 1; w <= 16; w++) { int ev = envs_per_sm(w); if (ev > best_env) { best_env = ev; best = w; } }
  // Warps of one block are kept in lock-step with block barriers (every sub-step and every Newton iteration) so that
  // they execute the same code at the same time and share instruction-cache lines; measured on B200 (A1, 4096 envs):
  // no barriers 417k, 2 blocks x 7 warps 564k, 1 block x 14 warps 633k env-steps/s -> take the largest block.
  h->sync_substeps = 1;
  for (int w = 16; w >= 1; w--) {
    if (envs_per_sm(w) == best_env) { best = w; break; }
  }
  {
    // A batch that does not even give every SM one full block is spread over all SMs with smaller blocks: one block per SM
    // beats half of the SMs running 14-15 warps (measured, BASELINE config 4: Atlas.walk 1024 envs as 147 blocks of 7
    // instead of 74 of 14, next to Talos.walk 1024: 634 k -> 834 k env-steps/s; 5 or 10 warps per block are both worse).
    int n_sm = 0;
    CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, h->device));
    // LOCOSIM_WPB=-1 keeps the largest block (MixedBatch: the lighter members of a mixed batch; 7-warp blocks of two different
    // kernels on every SM lost 0.7 ms per step to the cold start after an L2 flush, Atlas spread + Talos in full blocks did not).
    const bool keep = getenv("LOCOSIM_WPB") && atoi(getenv("LOCOSIM_WPB")) == -1;
    if (!keep && n_sm > 0 && (h->n_envs + best - 1) / best < n_sm) { int w = (h->n_envs + n_sm - 1) / n_sm; best = w < 1 ? 1 : (w < best ? w : best); }
  }
  if (getenv("LOCOSIM_WPB")) { int w = atoi(getenv("LOCOSIM_WPB")); if (w >= 1 && w <= 16 && w * per_env <= dev_max) best = w; }
  if (getenv("LOCOSIM_SYNC")) h->sync_substeps = atoi(getenv("LOCOSIM_SYNC"));
  h->regroup = 1;
  if (getenv("LOCOSIM_REGROUP")) h->regroup = atoi(getenv("LOCOSIM_REGROUP"));
  h->key_shift = C::RK4 ? 2 : 0;   // RK4: four solves per sub-step
  if (getenv("LOCOSIM_KEY_SHIFT")) h->key_shift = atoi(getenv("LOCOSIM_KEY_SHIFT"));
  if (getenv("LOCOSIM_KEY_MODE")) h->key_mode = atoi(getenv("LOCOSIM_KEY_MODE"));
  {
    // an MPR run holds its lock-step block about as long as `w` Newton iterations; key of a freshly reset env (+1; 0 = keep)
    int w = C::CONE == 1 ? 4 : 32, kr = C::CONE == 1 ? 1 : 0;       // (measured: profiles/README.md, round-2 knob sweeps)
    if (getenv("LOCOSIM_MPR_WEIGHT")) w = atoi(getenv("LOCOSIM_MPR_WEIGHT"));
    if (getenv("LOCOSIM_KEY_RESET")) kr = atoi(getenv("LOCOSIM_KEY_RESET")) + 1;
    h->key_mode = (h->key_mode & 255) | ((w & 255) << 8) | (kr << 16);
  }
  h->so.sync_iters = 16;    // lock-step group = the whole block
  if (getenv("LOCOSIM_SYNC_ITERS")) { int v = atoi(getenv("LOCOSIM_SYNC_ITERS")); h->so.sync_iters = v == 1 ? 16 : v; }
  if (getenv("LOCOSIM_GROUP")) h->so.sync_iters = atoi(getenv("LOCOSIM_GROUP"));
  if (h->so.sync_iters > 0 && (best + h->so.sync_iters - 1) / h->so.sync_iters > 15) h->so.sync_iters = 16;   // 15 named barriers
  h->so.sync_phases = 32;   // one more barrier where the warps enter the solver (measured +2%)
  if (getenv("LOCOSIM_SYNC_PHASES")) h->so.sync_phases = atoi(getenv("LOCOSIM_SYNC_PHASES"));
  if (h->so.sync_iters < best) h->so.sync_phases &= 63;   // block barriers inside the Newton loop need whole-block lock-step
  h->wpb = best;
  h->smem = best * per_env;
  { int dbg = getenv("LOCOSIM_DEBUG") ? atoi(getenv("LOCOSIM_DEBUG")) : 0; CK(cudaMemcpyToSymbol(c_debug, &dbg, sizeof(int))); }
  // ---- optional TMA staging of the mid-phase pair table (LOCOSIM_STAGE=1; measured, see profiles/README.md) ----
  // Default on wherever the table fits next to the per-env working sets (A1, Talos; not the RK4 configurations): A1 4096 envs
  // 2.7825 -> 2.7653 ms per step (+0.6 %, two alternating runs each). LOCOSIM_STAGE=0 turns it off.
  if (!getenv("LOCOSIM_STAGE") || atoi(getenv("LOCOSIM_STAGE")) > 0) {
    // the PRIMITIVE part of the table (the lean loop over the convex part streams the model's table from L2)
    const int np = h->hm.np_prim;
    const int bytes = (2 * np * 4 + 15) & ~15;                   // np packed pairs directly followed by np bounds
    if (np > 0 && h->smem + bytes + 16 <= dev_max) {
      CK(cudaMalloc((void**)&h->d_stage, bytes));
      CK(cudaMemset(h->d_stage, 0, bytes));
      CK(cudaMemcpy(h->d_stage, h->dm.pair_packed, 4 * np, cudaMemcpyDeviceToDevice));
      CK(cudaMemcpy(h->d_stage + 4 * np, h->dm.pair_bound, 4 * np, cudaMemcpyDeviceToDevice));
      h->dm.pk_n = np;
      CK(cudaMemcpyToSymbol(c_models, &h->dm, sizeof(DevModel), sizeof(DevModel) * h->slot));
      h->stage_bytes = bytes;
    }
  }
  // (a property of the FUNCTION, shared by every handle of this instantiation whatever its block size: always the device maximum)
  CK(cudaFuncSetAttribute(step_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, dev_max));
  return 0;
}

extern "C" {

const char* locosim_last_error(const locosim_t* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int locosim_create(const int32_t* mi, int nmi, const double* mr, int nmr, const int32_t* ti, int nti, const double* tr,
                   int ntr, int n_envs, int device, uint64_t seed, int64_t env_off, locosim_t** out) {
  locosim_handle* h = new locosim_handle();
  *out = nullptr;
  auto fail = [&](const std::string& msg) { g_create_error = msg; delete h; return 1; };
  if (n_envs <= 0) return fail("n_envs must be positive");
  std::string err = parse_model(h->hm, mi, nmi, mr, nmr);
  if (!err.empty()) return fail(err);
  err = parse_task(h->ht, h->hm.nu, h->hm.nv, h->hm.ng, ti, nti, tr, ntr);
  if (!err.empty()) return fail(err);
  if (h->hm.nu > h->hm.nv) return fail("nu > nv");
  h->device = device; h->n_envs = n_envs; h->seed = seed; h->env_off = env_off;
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) return fail(std::string("no CUDA device: ") + cudaGetErrorString(ce));
  if (device < 0 || device >= ndev) return fail("bad device index");
  ce = cudaSetDevice(device);
  if (ce != cudaSuccess) return fail(cudaGetErrorString(ce));
  if (cfg_fits<CfgEllEuler>(h->hm, h->ht)) h->cfg = 0;
  else if (cfg_fits<CfgPyrEuler>(h->hm, h->ht)) h->cfg = 1;
  else if (cfg_fits<CfgPyrRK4>(h->hm, h->ht)) h->cfg = 2;
  else if (cfg_fits<CfgPyrEuler29>(h->hm, h->ht)) h->cfg = 3;
  else return fail("no compiled configuration fits this model (cone / integrator / sizes: see locosim_config.h)");
  h->so.tolerance = 1e-5f; h->so.ls_tolerance = 0.1f; h->so.ls_iter = 16;
  h->so.max_iter = h->hm.iterations < 20 ? h->hm.iterations : 20;
  auto up = [&](void** dst, const void* src, size_t bytes) -> bool {
    if (cudaMalloc(dst, bytes ? bytes : 4) != cudaSuccess) return false;
    if (bytes && cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return false;
    return true;
  };
  bool ok = up((void**)&h->d_mints, h->hm.ints.data(), h->hm.ints.size() * 4) &&
            up((void**)&h->d_mreals, h->hm.reals.data(), h->hm.reals.size() * 4) &&
            up((void**)&h->d_tints, h->ht.ints.data(), h->ht.ints.size() * 4) &&
            up((void**)&h->d_treals, h->ht.reals.data(), h->ht.reals.size() * 4);
  size_t N = (size_t)n_envs, nv = (size_t)h->hm.nv;
  ok = ok && cudaMalloc((void**)&h->st.qpos, N * nv * 4) == cudaSuccess && cudaMalloc((void**)&h->st.qvel, N * nv * 4) == cudaSuccess &&
       cudaMalloc((void**)&h->st.ws, N * nv * 4) == cudaSuccess && cudaMalloc((void**)&h->st.goal, N * 4 * 4) == cudaSuccess &&
       cudaMalloc((void**)&h->st.episode, N * 4) == cudaSuccess && cudaMalloc((void**)&h->st.counters, N * 32) == cudaSuccess &&
       cudaMalloc((void**)&h->st.dr_row, N * 4) == cudaSuccess && cudaMalloc((void**)&h->st.perm, N * 4) == cudaSuccess &&
       cudaMalloc((void**)&h->st.cursor, N * 4) == cudaSuccess &&
       up((void**)&h->d_pool, h->hm.default_row.data(), h->hm.default_row.size() * 4);
  if (!ok) { std::string m = std::string("device allocation failed: ") + cudaGetErrorString(cudaGetLastError()); locosim_destroy(h); g_create_error = m; return 1; }
  cudaMemset(h->st.qpos, 0, N * nv * 4); cudaMemset(h->st.qvel, 0, N * nv * 4); cudaMemset(h->st.ws, 0, N * nv * 4);
  cudaMemset(h->st.goal, 0, N * 16); cudaMemset(h->st.episode, 0, N * 4); cudaMemset(h->st.counters, 0, N * 32);
  cudaMemset(h->st.dr_row, 0, N * 4); cudaMemset(h->st.cursor, 0, N * 4);
  h->st.rot_angle = nullptr;
  { std::vector<int> id(N); for (size_t i = 0; i < N; i++) id[i] = (int)i; cudaMemcpy(h->st.perm, id.data(), N * 4, cudaMemcpyHostToDevice); }
  h->st.pool = h->d_pool; h->st.pool_K = 1;
  bind_model(h->dm, h->hm, h->d_mints, h->d_mreals);
  for (int k = 0; k < LS_MAX_SLOTS && h->slot < 0; k++) if (!g_slot_used[k]) { h->slot = k; g_slot_used[k] = true; }
  if (h->slot < 0) { locosim_destroy(h); g_create_error = "too many live locosim handles (max 16 per process)"; return 1; }
  if (cudaMemcpyToSymbol(c_models, &h->dm, sizeof(DevModel), sizeof(DevModel) * h->slot) != cudaSuccess) {
    g_create_error = std::string("cudaMemcpyToSymbol: ") + cudaGetErrorString(cudaGetLastError()); locosim_destroy(h); return 1;
  }
  bind_task(h->dt, h->ht, h->hm.nu, h->d_tints, h->d_treals);
  h->dt.po_user = h->hm.po[12];
  int rc = 0;
  switch (h->cfg) {
    case 0: rc = setup_cfg<CfgEllEuler>(h); break;
    case 1: rc = setup_cfg<CfgPyrEuler>(h); break;
    case 3: rc = setup_cfg<CfgPyrEuler29>(h); break;
    default: rc = setup_cfg<CfgPyrRK4>(h); break;
  }
  if (rc) { g_create_error = h->err; locosim_destroy(h); return 1; }
  *out = h;
  return 0;
}

void locosim_destroy(locosim_t* h) {
  if (!h) return;
  if (h->slot >= 0) g_slot_used[h->slot] = false;
  cudaFree(h->d_mints); cudaFree(h->d_mreals); cudaFree(h->d_tints); cudaFree(h->d_treals);
  cudaFree(h->st.qpos); cudaFree(h->st.qvel); cudaFree(h->st.ws); cudaFree(h->st.goal); cudaFree(h->st.episode);
  cudaFree(h->st.counters); cudaFree(h->st.dr_row); cudaFree(h->st.perm); cudaFree(h->st.cursor); cudaFree(h->d_pool);
  cudaFree(h->d_stage);
  delete h;
}

int locosim_num_envs(const locosim_t* h) { return h->n_envs; }
int locosim_obs_dim(const locosim_t* h) { return h->ht.obs_dim; }
int locosim_action_dim(const locosim_t* h) { return h->hm.nu; }
int locosim_nq(const locosim_t* h) { return h->hm.nv; }

int locosim_set_solver(locosim_t* h, float tol, float ls_tol, int max_iter, int ls_iter) {
  if (tol <= 0 || ls_tol <= 0 || max_iter < 1 || ls_iter < 1) { h->err = "bad solver options"; return 1; }
  h->so.tolerance = tol; h->so.ls_tolerance = ls_tol; h->so.max_iter = max_iter; h->so.ls_iter = ls_iter;
  return 0;
}

int locosim_reset_rows(locosim_t* h, const uint8_t* d_mask, const int32_t* d_traj, const int32_t* d_step,
                       const int32_t* d_pool_row, float* d_obs, void* stream) {
  CK(cudaSetDevice(h->device));
  int threads = 128, blocks = (h->n_envs * 32 + threads - 1) / threads;
  reset_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(h->slot, h->dt, h->st, d_mask, d_traj, d_step, d_pool_row, d_obs,
                                                             h->n_envs, h->seed, h->env_off);
  h->st.rot_angle = nullptr;        // pinned rotation angles apply to one reset call
  CK(cudaGetLastError());
  return 0;
}
int locosim_set_reset_rotation(locosim_t* h, const float* d_angle) { h->st.rot_angle = d_angle; return 0; }
int locosim_get_cursor(locosim_t* h, int32_t* d_out, void* stream) {
  CK(cudaMemcpyAsync(d_out, h->st.cursor, (size_t)h->n_envs * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

// create_dataset() on the device (base.py:278-312 -> Trajectory.create_dataset utils/trajectory.py:104-151): states /
// next_states = the observation layout of consecutive samples of every trajectory of the reset table (already in HBM),
// last = 1 on the final transition of each trajectory; absorbing is all zero in the reference and is not materialised.
__global__ void dataset_kernel(int ms, DevTask t, float* __restrict__ states, float* __restrict__ next_states,
                               float* __restrict__ last) {
  const DevModel& m = c_models[ms];
  const int nv = m.nv, ncol = 2 * nv + t.n_goal, D = t.obs_dim, per = t.traj_len - 1;
  const long n_rows = (long)t.n_traj * per;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows * D; i += (long)gridDim.x * blockDim.x) {
    const long r = i / D;
    const int k = (int)(i - r * D), tj = (int)(r / per), sp = (int)(r - (long)tj * per);
    const float* row = t.table + ((size_t)tj * t.traj_len + sp) * ncol;
    const int ty = t.obs_src_type[k], idx = t.obs_src_idx[k];
    float a, b;
    if (ty == LS_OBS_QPOS) { a = row[idx]; b = row[ncol + idx]; }
    else if (ty == LS_OBS_QVEL) { a = row[nv + idx]; b = row[ncol + nv + idx]; }
    else if (ty == LS_OBS_GOAL) { a = row[2 * nv + idx]; b = row[ncol + 2 * nv + idx]; }
    else { a = 0.0f; b = 0.0f; }                      // foot forces / per-model features are not trajectory data
    states[i] = a; next_states[i] = b;
    if (k == 0 && last) last[r] = sp == per - 1 ? 1.0f : 0.0f;
  }
}
int locosim_dataset_rows(const locosim_t* h) { return h->ht.n_traj * (h->ht.traj_len - 1); }
int locosim_create_dataset(locosim_t* h, float* d_states, float* d_next_states, float* d_last, void* stream) {
  if (!d_states || !d_next_states) { h->err = "null buffer"; return 1; }
  CK(cudaSetDevice(h->device));
  dataset_kernel<<<296, 256, 0, (cudaStream_t)stream>>>(h->slot, h->dt, d_states, d_next_states, d_last);
  CK(cudaGetLastError());
  return 0;
}
int locosim_reset(locosim_t* h, const uint8_t* d_mask, const int32_t* d_traj, const int32_t* d_step, float* d_obs,
                  void* stream) {
  return locosim_reset_rows(h, d_mask, d_traj, d_step, nullptr, d_obs, stream);
}

int locosim_step(locosim_t* h, const float* a, float* o, float* r, uint8_t* d, float* no, int auto_reset, void* stream) {
  if (!a || !o || !r || !d) { h->err = "null buffer"; return 1; }
  CK(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  switch (h->cfg) {
    case 0: return launch_step<CfgEllEuler>(h, a, o, r, d, no, auto_reset, s);
    case 1: return launch_step<CfgPyrEuler>(h, a, o, r, d, no, auto_reset, s);
    case 3: return launch_step<CfgPyrEuler29>(h, a, o, r, d, no, auto_reset, s);
    default: return launch_step<CfgPyrRK4>(h, a, o, r, d, no, auto_reset, s);
  }
}

int locosim_get_state(locosim_t* h, float* q, float* v, float* w, void* stream) {
  size_t bytes = (size_t)h->n_envs * h->hm.nv * 4;
  cudaStream_t s = (cudaStream_t)stream;
  if (q) CK(cudaMemcpyAsync(q, h->st.qpos, bytes, cudaMemcpyDeviceToDevice, s));
  if (v) CK(cudaMemcpyAsync(v, h->st.qvel, bytes, cudaMemcpyDeviceToDevice, s));
  if (w) CK(cudaMemcpyAsync(w, h->st.ws, bytes, cudaMemcpyDeviceToDevice, s));
  return 0;
}
int locosim_set_state(locosim_t* h, const float* q, const float* v, const float* w, void* stream) {
  size_t bytes = (size_t)h->n_envs * h->hm.nv * 4;
  cudaStream_t s = (cudaStream_t)stream;
  if (q) CK(cudaMemcpyAsync(h->st.qpos, q, bytes, cudaMemcpyDeviceToDevice, s));
  if (v) CK(cudaMemcpyAsync(h->st.qvel, v, bytes, cudaMemcpyDeviceToDevice, s));
  if (w) CK(cudaMemcpyAsync(h->st.ws, w, bytes, cudaMemcpyDeviceToDevice, s));
  else CK(cudaMemsetAsync(h->st.ws, 0, bytes, s));
  return 0;
}
int locosim_set_goal(locosim_t* h, const float* g, void* stream) {
  if (!g) { h->err = "null buffer"; return 1; }
  CK(cudaMemcpyAsync(h->st.goal, g, (size_t)h->n_envs * 16, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}
int locosim_get_counters(locosim_t* h, int32_t* d_out, void* stream) {
  CK(cudaMemcpyAsync(d_out, h->st.counters, (size_t)h->n_envs * 32, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}
int locosim_param_pool_row_len(const locosim_t* h) { return h->hm.pool_P; }

int locosim_get_param_rows(locosim_t* h, int32_t* d_out, void* stream) {
  CK(cudaMemcpyAsync(d_out, h->st.dr_row, (size_t)h->n_envs * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

int locosim_set_param_pool(locosim_t* h, const double* pool, int n_rows, int row_len) {
  if (!pool || n_rows < 1 || row_len != h->hm.pool_P) { h->err = "bad parameter pool (row_len must equal locosim_param_pool_row_len)"; return 1; }
  CK(cudaSetDevice(h->device));
  std::vector<float> f((size_t)n_rows * row_len);
  bool damp = false;
  for (size_t i = 0; i < f.size(); i++) {
    f[i] = (float)pool[i];
    if (!(f[i] == f[i]) || f[i] > 1e30f || f[i] < -1e30f) { h->err = "non-finite value in parameter pool"; return 1; }
  }
  for (int r = 0; r < n_rows; r++)
    for (int d = 0; d < h->hm.nv; d++) if (f[(size_t)r * row_len + h->hm.po[0] + d] > 0) damp = true;
  float* d_new = nullptr;
  CK(cudaMalloc((void**)&d_new, f.size() * 4));
  CK(cudaMemcpy(d_new, f.data(), f.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaDeviceSynchronize());
  cudaFree(h->d_pool);
  h->d_pool = d_new; h->st.pool = d_new; h->st.pool_K = n_rows;
  // a dof gets a (static) frictionloss row as soon as ANY pool row gives it a positive frictionloss; in rows where
  // its value is 0 the constraint row is inert (force interval [-0, 0])
  std::vector<int> frow(h->hm.nv, -1);
  int nfric = 0;
  for (int d = 0; d < h->hm.nv; d++) {
    bool any = h->hm.default_row[h->hm.po[1] + d] > 0;
    for (int r = 0; r < n_rows && !any; r++) any = f[(size_t)r * row_len + h->hm.po[1] + d] > 0;
    if (any) frow[d] = nfric++;
  }
  CK(cudaMemcpy((void*)h->dm.dof_frow, frow.data(), frow.size() * 4, cudaMemcpyHostToDevice));
  h->dm.nfric = nfric;
  if (damp) h->dm.has_damping = 1;
  CK(cudaMemcpyToSymbol(c_models, &h->dm, sizeof(DevModel), sizeof(DevModel) * h->slot));
  CK(cudaMemset(h->st.dr_row, 0, (size_t)h->n_envs * 4));
  return 0;
}

int locosim_kernels_per_step(const locosim_t* h) { return h->regroup ? 2 : 1; }
int locosim_debug_counters(unsigned long long* out8) {
  return cudaMemcpyFromSymbol(out8, g_dbg, 64) == cudaSuccess ? 0 : 1;
}

// Measured non-tensor FP32 peak (the denominator of the FP32 utilisation bench.py reports next to the HBM roofline):
// 8 independent FMA chains per thread, 2 x 256 threads per SM slot, all SMs, best of 5 launches.
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* out, int iters) {
  float a[8];
#pragma unroll
  for (int k = 0; k < 8; k++) a[k] = 1.0f + 1e-3f * (float)(threadIdx.x + k);
  const float b = 0.9999f, c = 1e-4f;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int k = 0; k < 8; k++) a[k] = fmaf(a[k], b, c);
    }
  }
  float s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += a[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int locosim_measure_fp32_peak(int device, double* tflops_out) {
  if (!tflops_out) return 1;
  if (cudaSetDevice(device) != cudaSuccess) return 1;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const int blocks = sms * 8, threads = 256, iters = 4096;
  float* out = nullptr;
  if (cudaMalloc((void**)&out, (size_t)blocks * threads * 4) != cudaSuccess) return 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 6; rep++) {
    cudaEventRecord(e0);
    fp32_peak_kernel<<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(out); return 1; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 64.0 * (double)iters * (double)blocks * threads;
    if (rep > 0 && ms > 0) best = flops / (ms * 1e-3) / 1e12 > best ? flops / (ms * 1e-3) / 1e12 : best;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  *tflops_out = best;
  return 0;
}

int locosim_launch_info(const locosim_t* h, int* wpb, int* smem, int* blocks) {
  if (wpb) *wpb = h->wpb;
  if (smem) *smem = h->smem;
  if (blocks) *blocks = (h->n_envs + h->wpb - 1) / h->wpb;
  return 0;
}

}  // extern "C"
