"""Probe of BASELINE config 4 (Atlas.walk + Talos.walk with domain randomisation, 1024 + 1024 envs on one GPU): launch
geometry variants (warps per block per engine) -> per-step time over a long rollout.   (under gpurun)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("LOCO_MUJOCO_B200_FORCE_BUNDLED", "1")
import bench  # noqa: E402


def main():
    import torch
    dr = lambda robot: "domain_randomization_%s.yaml" % robot
    members = [("Atlas.walk", 1024, {"domain_randomization_config": dr("atlas")}),
               ("Talos.walk", 1024, {"domain_randomization_config": dr("talos")})]
    T = 260
    for label, wpbs, flush_on in [("default flush", None, True), ("default flush", None, True), ("default noflush", None, False)]:
        mem = members if wpbs is None else [(t, n, dict(kw, warps_per_block=w)) for (t, n, kw), w in zip(members, wpbs)]
        wl = bench.Workload("cfg4", mem, 0, 1, 0)
        dev = wl.dev
        wl.batch.reset()
        gen = torch.Generator(device=dev).manual_seed(1234)
        actions = [torch.rand((T, e.n_envs, e.action_dim), device=dev, generator=gen) * 2 - 1 for e in wl.engines]
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
        ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(T)]
        ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(T)]
        for k in range(T):
            if flush_on:
                flush.fill_(k & 0xff)
            ev0[k].record()
            wl.step([x[k] for x in actions])
            ev1[k].record()
        torch.cuda.synchronize()
        ms = [a.elapsed_time(b) for a, b in zip(ev0, ev1)]
        win = ["%.2f" % (sum(ms[i:i + 20]) / 20) for i in range(0, T, 20)]
        print(label, [e.launch_info()["warps_per_block"] for e in wl.engines], "ms/step per 20-step window:", " ".join(win), flush=True)
        del wl, flush
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
