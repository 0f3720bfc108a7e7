"""Batch sharding of the sampling path across the GPUs of one node.

Complexes are independent -- no operation of FullDPM.sample mixes the batch dimension -- so the batch is split
contiguously over the ranks, every rank runs the full T-step loop on its own complexes with replicated weights and
NO collective inside the loop, and one all_gather of the finished structures closes the run.  One process per GPU,
torch.distributed (NCCL on GPUs; the host logic is backend agnostic and is tested with gloo on CPU).
"""
import torch
import torch.distributed as dist


def shard_bounds(n, world_size, rank):
    """Contiguous split of n items: the first n % world_size ranks get one extra."""
    base, extra = divmod(n, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(batch, world_size, rank):
    """Slice every tensor of `batch` (dict name -> tensor with the complexes on dim 0) to this rank's share."""
    n = next(iter(batch.values())).shape[0]
    a, b = shard_bounds(n, world_size, rank)
    return {k: v[a:b].contiguous() for k, v in batch.items()}


def gather_results(local, n_total, group=None):
    """all_gather of per-complex results (tensors with the local complexes on dim 0) back into batch order.

    Shards may differ by one complex; they are padded to the largest shard for the collective and trimmed after.
    Returns a list of tensors with n_total complexes on dim 0, identical on every rank."""
    world = dist.get_world_size(group)
    cap = -(-n_total // world)
    out = []
    for t in local:
        pad = torch.zeros((cap,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
        buf = torch.empty((world * cap,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(buf, pad, group=group)
        parts = []
        for r in range(world):
            a, b = shard_bounds(n_total, world, r)
            parts.append(buf[r * cap:r * cap + (b - a)])
        out.append(torch.cat(parts, 0))
    return out


def sample_sharded(model, v, p, s, res_feat, pair_feat, mask_generate, mask_res, group=None, **kw):
    """FullDPM.sample on this rank's share of the batch + one gather of traj[0] = (v, p, s).

    Every rank passes the FULL batch (host or device tensors); returns (v, p, s) for all complexes on every rank."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = v.shape[0]
    mine = shard_batch(dict(v=v, p=p, s=s, res_feat=res_feat, pair_feat=pair_feat, mask_generate=mask_generate,
                            mask_res=mask_res), world, rank)
    dev = next(model.parameters()).device
    mine = {k: t.to(dev) for k, t in mine.items()}
    traj = model.sample(mine['v'], mine['p'], mine['s'], mine['res_feat'], mine['pair_feat'], mine['mask_generate'],
                        mine['mask_res'], **kw)
    return gather_results([traj[0][0], traj[0][1], traj[0][2]], n, group)
