"""Host logic of the multi-GPU path with two gloo ranks on CPU: contiguous batch sharding,
no collective inside the loop, one gather of the finished structures."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ab_opt_b200 import sharding


def test_shard_bounds_cover_the_batch():
    for n in (1, 7, 64, 65, 256):
        for w in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_is_bit_exact():
    g = torch.Generator().manual_seed(1)
    v, p = torch.randn(5, 7, 3, generator=g), torch.randn(5, 7, 3, generator=g) * 1e3
    s = torch.randint(-2 ** 62, 2 ** 62, (5, 7), generator=g)          # int64 travels as two float32 words, bit for bit
    packed = sharding.pack_results([v, p, s])
    assert packed.dtype == torch.float32 and packed.shape == (5, 7 * 3 + 7 * 3 + 7 * 2)
    v2, p2, s2 = sharding.unpack_results(packed, [v, p, s])
    assert torch.equal(v, v2) and torch.equal(p, p2) and torch.equal(s, s2) and s2.dtype == torch.int64
    e = sharding.unpack_results(packed[:0], [v, p, s])
    assert e[2].shape == (0, 7)


class _FakeDPM(torch.nn.Module):
    """Stands in for FullDPM on CPU: a per-complex deterministic 'sample' (no batch mixing, like the real loop)."""

    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.ones(1))

    calls = []

    def sample(self, v, p, s, res_feat, pair_feat, mask_generate, mask_res, **kw):
        _FakeDPM.calls.append(kw)
        g = mask_generate[..., None]
        v0 = torch.where(g, v + res_feat[..., :3].tanh(), v)
        p0 = torch.where(g, p + pair_feat.mean(dim=(2, 3))[..., None], p)
        s0 = torch.where(mask_generate, (s + 1) % 20, s)
        return {0: (v0, p0, s0)}


def _worker(rank, world, port, n, L, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)           # every rank builds the same full batch
    batch = dict(v=torch.randn(n, L, 3, generator=g), p=torch.randn(n, L, 3, generator=g), s=torch.randint(0, 20, (n, L), generator=g),
                 res_feat=torch.randn(n, L, 128, generator=g), pair_feat=torch.randn(n, L, L, 64, generator=g),
                 mask_generate=torch.rand(n, L, generator=g) < 0.3, mask_res=torch.ones(n, L, dtype=torch.bool))
    model = _FakeDPM()
    mine = sharding.shard_batch(batch, world, rank)
    a, b = sharding.shard_bounds(n, world, rank)
    assert mine['v'].shape[0] == b - a and torch.equal(mine['s'], batch['s'][a:b])
    got = sharding.sample_sharded(model, **batch)
    if b > a:
        kw = _FakeDPM.calls[-1]                    # the loop call carries the shard's place in the global batch and a common seed
        assert kw['batch_offset'] == a and kw['batch_total'] == n and isinstance(kw['seed'], int)
        seed = kw['seed']
    else:
        assert not _FakeDPM.calls                  # an empty shard (fewer complexes than ranks) never enters the loop
        seed = None
    seeds = [None, None]
    dist.all_gather_object(seeds, seed)
    assert seeds[0] == seeds[1] or None in seeds
    ref = model.sample(**batch)[0]
    ok = all(torch.equal(x, y) for x, y in zip(got, ref))
    ret[rank] = bool(ok) and got[0].shape[0] == n
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather():
    with socket.socket() as sk:
        sk.bind(('127.0.0.1', 0))
        port = sk.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, 5, 12, ret), nprocs=2, join=True)      # 5 complexes over 2 ranks: uneven shards
    assert ret[0] and ret[1]


def test_two_rank_gloo_fewer_complexes_than_ranks():
    with socket.socket() as sk:
        sk.bind(('127.0.0.1', 0))
        port = sk.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, 1, 9, ret), nprocs=2, join=True)       # 1 complex over 2 ranks: rank 1's shard is empty
    assert ret[0] and ret[1]
