"""Batch sharding of the sampling path across the GPUs of one node.

Complexes are independent -- no operation of FullDPM.sample mixes the batch dimension -- so the batch is split
contiguously over the ranks, every rank runs the full T-step loop on its own complexes with replicated weights and
NO collective inside the loop, and ONE all_gather of the finished structures (v, p, s packed into one buffer) closes the
run.  One process per GPU, torch.distributed (NCCL on GPUs; the host logic is backend agnostic and is tested with gloo
on CPU).

A sharded run reproduces the single-device run of the whole batch: every rank passes its offset in the global batch, the
in-kernel Philox counters are global residue rows (abopt_model_set_batch_offset) and the seed is rank 0's.
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


def pack_results(local):
    """Per-complex result tensors (complexes on dim 0; float32 or int64) -> one (n_local, width) float32 buffer.
    int64 tensors travel as two float32 words each (a bit-exact view, not a conversion)."""
    cols = []
    for t in local:
        per = 1
        for d in t.shape[1:]:
            per *= int(d)
        flat = t.reshape(t.shape[0], per).contiguous()      # (explicit width: an empty shard cannot infer -1)
        if flat.dtype == torch.int64:
            flat = flat.view(torch.float32)
        elif flat.dtype != torch.float32:
            raise TypeError(f'unsupported dtype {t.dtype}')
        cols.append(flat)
    return torch.cat(cols, 1).contiguous()


def unpack_results(packed, like):
    """Inverse of pack_results for tensors shaped / typed like `like` (any number of complexes on dim 0)."""
    out, c0 = [], 0
    n = packed.shape[0]
    for t in like:
        per = 1
        for d in t.shape[1:]:
            per *= int(d)
        w = per * (2 if t.dtype == torch.int64 else 1)
        part = packed[:, c0:c0 + w].contiguous()
        if t.dtype == torch.int64:
            part = part.view(torch.int64)
        out.append(part.reshape((n,) + tuple(t.shape[1:])))
        c0 += w
    return out


def gather_results(local, n_total, group=None):
    """ONE all_gather of the per-complex results (tensors with the local complexes on dim 0) back into batch order.

    Shards may differ by one complex; they are padded to the largest shard for the collective and trimmed after.
    Returns a list of tensors with n_total complexes on dim 0, identical on every rank."""
    world = dist.get_world_size(group)
    cap = -(-n_total // world)
    packed = pack_results(local)
    pad = torch.zeros((cap, packed.shape[1]), dtype=packed.dtype, device=packed.device)
    pad[:packed.shape[0]] = packed
    buf = torch.empty((world * cap, packed.shape[1]), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(buf, pad, group=group)
    parts = []
    for r in range(world):
        a, b = shard_bounds(n_total, world, r)
        parts.append(buf[r * cap:r * cap + (b - a)])
    return unpack_results(torch.cat(parts, 0), local)


def common_seed(model, group=None, seed=None, **kw):
    """The Philox seed every rank must share: `seed` if given, else rank 0's draw from torch's generator, broadcast (8 bytes,
    before the loop starts -- not a data-path collective).  None in parity mode (rng='torch')."""
    if seed is not None or kw.get('rng', getattr(model, 'rng', 'philox')) != 'philox':
        return seed
    dev = next(model.parameters()).device
    sd = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).to(dev)
    dist.broadcast(sd, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return int(sd.item())


def sample_shard(model, mine, offset, n_total, group=None, seed=None, **kw):
    """The per-rank half of sample_sharded for callers that already hold only their own complexes (`mine`: dict with v, p, s,
    res_feat, pair_feat, mask_generate, mask_res on the model's device; `offset` = index of mine[0] in the global batch of
    `n_total`): the full T-step loop on the shard, then the ONE packed gather.  Returns (v, p, s) of all complexes."""
    seed = common_seed(model, group, seed, **kw)
    if mine['v'].shape[0] == 0:
        # fewer complexes than ranks: this rank has nothing to sample, but it still takes part in the seed broadcast above
        # and in the gather (zero rows)
        return gather_results([mine['v'].float(), mine['p'].float(), mine['s'].long()], n_total, group), {}
    traj = model.sample(mine['v'], mine['p'], mine['s'], mine['res_feat'], mine['pair_feat'], mine['mask_generate'],
                        mine['mask_res'], seed=seed, batch_offset=offset, batch_total=n_total, **kw)
    return gather_results([traj[0][0], traj[0][1], traj[0][2]], n_total, group), traj


def sample_sharded(model, v, p, s, res_feat, pair_feat, mask_generate, mask_res, group=None, seed=None, **kw):
    """FullDPM.sample on this rank's share of the batch + one gather of traj[0] = (v, p, s).

    Every rank passes the FULL batch (host or device tensors); returns (v, p, s) for all complexes on every rank.  `seed`
    must be the same on every rank (default: rank 0's, see common_seed); with rng='torch' every rank must hold the same
    generator state."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = v.shape[0]
    a, _ = shard_bounds(n, world, rank)
    mine = shard_batch(dict(v=v, p=p, s=s, res_feat=res_feat, pair_feat=pair_feat, mask_generate=mask_generate,
                            mask_res=mask_res), world, rank)
    dev = next(model.parameters()).device
    mine = {k: t.to(dev) for k, t in mine.items()}
    return sample_shard(model, mine, a, n, group, seed, **kw)[0]
