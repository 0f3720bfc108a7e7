"""Two NCCL ranks on two GPUs (skipped on a single-GPU box): ab_opt_b200.sharding.sample_sharded = contiguous batch split, no
collective inside the loop, ONE packed all_gather at the end -- and the gathered result equals the single-GPU run of the whole
batch bit for bit in both RNG modes ("parity mode")."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret):
    import ab_opt_b200
    from ab_opt_b200 import sharding
    from oracle import weights
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dev = torch.device('cuda', rank)
    torch.cuda.set_device(dev)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    W = weights.make_state_dict(seed=3, num_layers=2, flavour='abdesign')
    model = ab_opt_b200.FullDPMAbDesign(128, 64, 100, eps_net_opt=dict(num_layers=2))
    model.load_state_dict(W, strict=True)
    model = model.to(dev).eval()
    n, L = 5, 72                                                  # uneven shards: 3 + 2 complexes
    inp = weights.synthetic_inputs(9, n, L, gen_slices=((10, 22), (40, 44)), ragged=True)
    args = [inp[k] for k in ('v', 'p', 's', 'res_feat', 'pair_feat', 'mask_generate', 'mask_res')]
    ok = True
    calls = []
    orig = dist.all_gather_into_tensor
    dist.all_gather_into_tensor = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    got = sharding.sample_sharded(model, *args, seed=77)
    dist.all_gather_into_tensor = orig
    ok &= len(calls) == 1                                         # one gather, nothing else on the data path
    whole = model.sample(*[x.to(dev) for x in args], seed=77)[0]
    ok &= all(torch.equal(g, w) for g, w in zip(got, whole))
    # parity mode: every rank holds the same generator state, draws the reference's full-batch tensors and uses its rows
    torch.manual_seed(5)
    got_t = sharding.sample_sharded(model, *args, rng='torch')
    torch.manual_seed(5)
    whole_t = model.sample(*[x.to(dev) for x in args], rng='torch')[0]
    ok &= all(torch.equal(g, w) for g, w in zip(got_t, whole_t))
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_nccl_sharded_sample_equals_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    with socket.socket() as sk:
        sk.bind(('127.0.0.1', 0))
        port = sk.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]
