/*
 * CPU ORACLE (test infrastructure, NOT the product).
 *
 * fp64 single-environment restatement of the reference's LocoEnv.step() hot path:
 *   mushroom_rl MuJoCo.step -> mujoco.mj_step(model, data, n_substeps)   (third-party, un-vendored:
 *   mujoco==2.3.7 pinned in /root/reference/pyproject.toml:9; call parameters fixed by
 *   /root/reference/loco_mujoco/environments/base.py:32-33,94-98,109-111)
 * restated from MuJoCo 2.3.7's published computation pipeline (mj_step = mj_forward + integrator) for the
 * feature set of the in-scope robots.  Pinned against the reference's golden rollouts
 * (/root/reference/tests/test_datasets/<task>.npy, copied to tests/golden/) in tests/test_oracle_golden.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 */
#ifndef LOCOSIM_REF_H
#define LOCOSIM_REF_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct RefSim RefSim;

RefSim* ref_create(const int* ints, int n_ints, const double* reals, int n_reals);
void ref_destroy(RefSim* s);
int ref_nv(const RefSim* s);
int ref_nu(const RefSim* s);

/* mj_resetData + write qpos/qvel (LocoEnv.reset -> set_sim_state, base.py:178-203,478-497) */
void ref_reset(RefSim* s, const double* qpos, const double* qvel);
void ref_get_state(const RefSim* s, double* qpos, double* qvel);
void ref_get_warmstart(const RefSim* s, double* qacc_warmstart);
void ref_set_warmstart(RefSim* s, const double* qacc_warmstart);

/* data.ctrl[:] = ctrl ; mujoco.mj_step(model, data, nsub) */
void ref_step(RefSim* s, const double* ctrl, int nsub);

/* diagnostics of the last mj_forward */
int ref_ncon(const RefSim* s);
int ref_nefc(const RefSim* s);
int ref_solver_iter(const RefSim* s);
void ref_get_qacc(const RefSim* s, double* qacc);
/* contact k: out[0]=dist, out[1..3]=pos, out[4..6]=normal, out[7]=geom1, out[8]=geom2, out[9]=dim; then force[6] */
void ref_get_contact(const RefSim* s, int k, double* out16);
/* per-geom-group normal-frame contact force sum like mj_contactForce over contacts between g1set and g2set */
void ref_forward(RefSim* s, const double* ctrl);   /* mj_forward only (no integration) */
void ref_get_M(const RefSim* s, double* M);         /* dense nv*nv */
void ref_get_bias(const RefSim* s, double* qfrc_bias);

/*
 * Batched environment loop used as the CPU baseline (bench.py): n independent envs, `nthreads` host threads,
 * random actions, auto-reset from the trajectory table; returns the number of env-steps executed.
 * task: see loco_mujoco_b200/task.py TaskSpec.pack() (same blob the CUDA engine takes).
 */
long ref_rollout(const int* ints, int n_ints, const double* reals, int n_reals,
                 const int* task_ints, int n_task_ints, const double* task_reals, int n_task_reals,
                 int n_envs, int n_steps, int nthreads, unsigned long long seed,
                 double* obs_out /* [n_envs, obs_dim] last obs, may be NULL */,
                 long* n_resets_out);
/* Same loop with the workers pinned one per allowed CPU (pin != 0) and the time they actually got measured:
 * stats_out[4] = {wall seconds, sum of worker CPU seconds, slowest worker's CPU seconds, CPUs pinned to}.
 * cpu/wall is the number of cores the host really granted (a cgroup quota or a busy neighbour shows up here). */
long ref_rollout_ex(const int* ints, int n_ints, const double* reals, int n_reals,
                    const int* task_ints, int n_task_ints, const double* task_reals, int n_task_reals,
                    int n_envs, int n_steps, int nthreads, unsigned long long seed, double* obs_out,
                    long* n_resets_out, int pin, double* stats_out);

/* single env step of the full LocoEnv.step contract (action in [-1,1]) -> obs, reward, absorbing */
typedef struct RefEnv RefEnv;
RefEnv* refenv_create(const int* ints, int n_ints, const double* reals, int n_reals,
                      const int* task_ints, int n_task_ints, const double* task_reals, int n_task_reals);
void refenv_destroy(RefEnv* e);
int refenv_obs_dim(const RefEnv* e);
/* user features observed through LS_OBS_PARAM (multi-model envs: the carried weight's mass) */
void refenv_set_user(RefEnv* e, const double* user);
/* setup_random_rot: rotation angle applied by the following refenv_reset_to calls (0 by default) */
void refenv_set_rotation(RefEnv* e, double angle);
/* trajectory cursor (traj * traj_len + sample) of the tracking reward */
int refenv_cursor(const RefEnv* e);
void refenv_reset_to(RefEnv* e, int traj_no, int step_no, double* obs);
void refenv_step(RefEnv* e, const double* action, double* obs, double* reward, int* absorbing);
RefSim* refenv_sim(RefEnv* e);

#ifdef __cplusplus
}
#endif
#endif
