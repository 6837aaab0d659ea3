"""
Multi-GPU plumbing: environments are independent, so the env axis is sharded across ranks, one process
per GPU (`torch.distributed`, NCCL over NVLink on the B200 box, gloo in the CPU tests). There is no collective on the
data path of `step()`; the only (optional) exchange is an all-gather of the rollout buffer for consumers that want the
global batch on every rank.
"""
import torch
import torch.distributed as dist


def shard_range(n_total, world_size, rank):
    """Contiguous, balanced [offset, offset+count) slice of the global env axis owned by `rank`."""
    base, rem = divmod(int(n_total), int(world_size))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def make_sharded(task_id, n_total, rank=None, world_size=None, device=None, seed=0, **kwargs):
    """LocoEnv for this rank's shard; `env_id_offset` keeps every env's random stream identical to the 1-GPU run."""
    from . import LocoEnv
    rank = dist.get_rank() if rank is None else rank
    world_size = dist.get_world_size() if world_size is None else world_size
    offset, count = shard_range(n_total, world_size, rank)
    device = device or ("cuda:%d" % (rank % max(1, torch.cuda.device_count())))
    return LocoEnv.make(task_id, num_envs=count, device=device, seed=seed, env_id_offset=offset, **kwargs)


def gather_rollout(obs, reward, done, group=None):
    """all-gather of the per-rank rollout buffer -> (obs[N_total, D], reward[N_total], done[N_total]) in global env
    order. Shards may differ by one env (shard_range), so shorter shards are padded for the collective."""
    world = dist.get_world_size(group)
    n = torch.tensor([obs.shape[0]], device=obs.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    nmax = max(sizes)
    buf = torch.zeros((nmax, obs.shape[1] + 2), device=obs.device, dtype=obs.dtype)
    buf[:obs.shape[0], :-2] = obs
    buf[:obs.shape[0], -2] = reward
    buf[:obs.shape[0], -1] = done.to(obs.dtype)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    cat = torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)
    return cat[:, :-2], cat[:, -2], cat[:, -1] > 0.5


def aggregate_throughput(local_units, local_seconds, device=None, group=None):
    """Whole-job throughput = units of all ranks / max-over-ranks time (the bench contract)."""
    t = torch.tensor([float(local_seconds)], dtype=torch.float64, device=device)
    u = torch.tensor([float(local_units)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(u, op=dist.ReduceOp.SUM, group=group)
    return float(u.item()) / float(t.item()), float(t.item())
