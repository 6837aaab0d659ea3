"""Done-flag flip rate and long-horizon divergence of the fp32 engine against the fp64 oracle (GPU box).

SURVEY.md section 7(5): `done` thresholds are evaluated identically (strict < / >), so an env whose terminating quantity is
within fp32 error of its threshold can legitimately flip. This tool measures how often, and how fast free-running fp32 and
fp64 rollouts separate, at random-action (U(-1,1)) steady state:

  (A) one-step comparison from IDENTICAL states: before every GPU step the batch state (qpos, qvel, warm start, fp32 values)
      is handed to one oracle env per GPU env, both take the same action; compared: done flag, observation.
  (B) free-running: same reset rows, same action sequence, no re-synchronisation; reported: the step at which
      |obs_gpu - obs_oracle| first exceeds 1e-2 and the episode lengths of both.

    python tools/flip_rate.py [--envs 256] [--steps 60] [--out gpurun_out/flip_rate.json]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("LOCO_MUJOCO_B200_FORCE_BUNDLED", "1")
from loco_mujoco_b200 import LocoEnv, modelpack  # noqa: E402
import oracle_binding  # noqa: E402


def near_threshold(env, obs, eps):
    for key, lo, hi in env._has_fallen_terms():
        v = obs[env.get_obs_idx(key)[0]]
        if abs(v - lo) < eps or abs(v - hi) < eps:
            return True
    return False


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--envs", type=int, default=256)
    p.add_argument("--steps", type=int, default=60)
    p.add_argument("--tasks", nargs="*", default=["UnitreeA1.simple", "HumanoidTorque.run", "Atlas.walk", "Talos.walk"])
    p.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "flip_rate.json"))
    a = p.parse_args()
    oracle = oracle_binding.load(os.path.join(ROOT, "oracle", "liblocosim_ref.so"))
    report = {}
    for task in a.tasks:
        n = a.envs
        env = LocoEnv.make(task + ".real", debug=True, num_envs=n, seed=7, copy_outputs=False)
        eng = env._get_engine()
        mb, tb = modelpack.pack(env._model), env.task_spec().pack()
        oes = [oracle.env(mb, tb) for _ in range(n)]
        for oe in oes:
            oe.reset_to(0, 0)
        env.reset()
        gen = torch.Generator(device="cpu").manual_seed(11)
        # pre-roll to the stationary regime
        for _ in range(30):
            env.step((torch.rand((n, eng.action_dim), generator=gen) * 2 - 1).cuda())
        # ---- (A) per-step comparison from identical states ----
        cmp = dict(compared=0, done_gpu=0, done_oracle=0, mismatch=0, mismatch_near_threshold_1e3=0, mismatch_near_threshold_1e4=0)
        errs, rels = [], []
        for k in range(a.steps):
            act = torch.rand((n, eng.action_dim), generator=gen) * 2 - 1
            q, v, w = [x.double().cpu().numpy() for x in eng.get_state()]
            obs, rew, done, nxt = eng.step(act.cuda(), auto_reset=True)
            obs, done = obs.double().cpu().numpy(), done.cpu().numpy().astype(bool)
            for i, oe in enumerate(oes):
                oe.set_state(q[i], v[i], w[i])
                o, r, d = oe.step(act[i].double().numpy())
                cmp["compared"] += 1
                cmp["done_gpu"] += int(done[i])
                cmp["done_oracle"] += int(d)
                errs.append(float(np.abs(o - obs[i]).max()))
                rels.append(float((np.abs(o - obs[i]) / (1.0 + np.abs(o))).max()))
                if d != done[i]:
                    cmp["mismatch"] += 1
                    cmp["mismatch_near_threshold_1e3"] += int(near_threshold(env, o, 1e-3))
                    cmp["mismatch_near_threshold_1e4"] += int(near_threshold(env, o, 1e-4))
        errs = np.array(errs)
        cmp["flip_rate_per_env_step"] = cmp["mismatch"] / cmp["compared"]
        cmp["obs_err_p50_p99_max"] = [float(np.percentile(errs, 50)), float(np.percentile(errs, 99)), float(errs.max())]
        rels = np.array(rels)
        cmp["obs_err_relative_p50_p99_max"] = [float(np.percentile(rels, 50)), float(np.percentile(rels, 99)), float(rels.max())]
        cmp["note"] = "relative = max_k |d_k| / (1 + |obs_k|); random U(-1,1) torques drive joint velocities to O(100) rad/s"
        # ---- (B) free-running from the same reset rows ----
        rng = np.random.RandomState(3)
        tr = rng.randint(0, env.trajectories.number_of_trajectories, n).astype(np.int32)
        st = rng.randint(0, env.trajectories.trajectory_length, n).astype(np.int32)
        eng.reset(traj_no=torch.tensor(tr, device=eng.device), step_no=torch.tensor(st, device=eng.device))
        for i, oe in enumerate(oes):
            oe.reset_to(int(tr[i]), int(st[i]))
        alive_g, alive_o = np.ones(n, bool), np.ones(n, bool)
        len_g, len_o = np.zeros(n, int), np.zeros(n, int)
        sep = np.full(n, -1)
        for k in range(a.steps):
            act = torch.rand((n, eng.action_dim), generator=gen) * 2 - 1
            obs, rew, done, nxt = eng.step(act.cuda(), auto_reset=False)
            obs, done = obs.double().cpu().numpy(), done.cpu().numpy().astype(bool)
            for i, oe in enumerate(oes):
                if alive_o[i]:
                    o, r, d = oe.step(act[i].double().numpy())
                    len_o[i] += 1
                    if alive_g[i] and sep[i] < 0 and np.abs(o - obs[i]).max() > 1e-2:
                        sep[i] = k
                    if d:
                        alive_o[i] = False
                if alive_g[i]:
                    len_g[i] += 1
                    if done[i]:
                        alive_g[i] = False
        free = dict(mean_episode_len_gpu=float(len_g.mean()), mean_episode_len_oracle=float(len_o.mean()),
                    same_episode_len_frac=float((len_g == len_o).mean()),
                    episode_len_diff_le1_frac=float((np.abs(len_g - len_o) <= 1).mean()),
                    separated_1e2_before_end_frac=float((sep >= 0).mean()),
                    median_separation_step=float(np.median(sep[sep >= 0])) if (sep >= 0).any() else None)
        report[task] = {"one_step_from_identical_state": cmp, "free_running": free, "envs": n, "steps": a.steps}
        print(task, json.dumps(report[task]), flush=True)
        for oe in oes:
            oe.close()
        del env, eng
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(report, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
