"""GPU parity of the steps after the loop: reconstruct_backbone_partially and the RMSD ranking, through
ab_opt_b200.post (C ABI: abopt_reconstruct_backbone_partially / abopt_pairwise_rmsd / abopt_rank_commoness) against the
reference fixture tests/golden/post_loop.npz and the CPU oracle."""
import os

import numpy as np
import pytest
import torch

import ab_opt_b200
from ab_opt_b200 import post as P
from oracle import post as OP, pair_embed as PE, geometry as G

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def assert_atoms_close(got, want):
    """N, CA, C: a rigid placement, a few ulps of a ~50 A coordinate.  O: psi enters through acos of a clamped cosine
    (geometry.py:269) whose derivative reaches 707 at the clamp, so fp32 noise in the cosine moves O (2.4 A from CA) by up to
    ~1e-4 A for the ~0.1 % of residues with a near-planar N-CA-C-N; everything else (zero padding, context atoms) is exact."""
    torch.testing.assert_close(got[:, :, :3], want[:, :, :3], rtol=2e-6, atol=2e-5)
    torch.testing.assert_close(got[:, :, 3], want[:, :, 3], rtol=2e-6, atol=5e-4)
    assert (got[:, :, 3] - want[:, :, 3]).abs().median() <= 2e-5
    assert torch.equal(got[:, :, 4:], want[:, :, 4:])


@pytest.fixture(scope='module')
def gold(golden_dir):
    d = np.load(os.path.join(golden_dir, 'post_loop.npz'))
    return {k: torch.from_numpy(d[k]) if d[k].ndim else d[k].item() for k in d.files}


def test_reconstruct_against_reference_fixture(gold):
    g = gold
    inp = PE.synthetic_complex(5, 2, 20)
    args = (inp['pos_atoms'], G.so3_exp(g['v']), g['t'], g['aa'], inp['chain_nb'], inp['res_nb'], inp['mask_atoms'], g['mask_recons'])
    # CUDA tensors in -> CUDA tensors out
    pos_new, mask_new = P.reconstruct_backbone_partially(*[x.to(DEV) for x in args], bb_table=g['bb_table'], o_table=g['o_table'])
    assert pos_new.is_cuda and mask_new.dtype == torch.bool
    assert_atoms_close(pos_new.cpu(), g['pos_new'])
    assert torch.equal(mask_new.cpu(), g['mask_new'])
    # CPU tensors in (as the reference's runner passes them) -> computed on the GPU, returned on the CPU
    P.set_backbone_tables(g['bb_table'], g['o_table'])
    pos_cpu, mask_cpu = P.reconstruct_backbone_partially(*args)
    assert not pos_cpu.is_cuda and torch.equal(pos_cpu, pos_new.cpu()) and torch.equal(mask_cpu, mask_new.cpu())


@pytest.mark.parametrize('N,L,A', [(3, 37, 15), (1, 1, 4), (101, 64, 15), (64, 256, 15)])
def test_reconstruct_against_oracle(gold, N, L, A):
    """All residues rebuilt / none / a stretch; a whole 101-frame trajectory in one launch; the C2 batch."""
    inp = PE.synthetic_complex(60 + L, N, L, num_atoms_in=A)
    gen = torch.Generator().manual_seed(N)
    R = G.so3_exp(torch.randn(N, L, 3, generator=gen))
    t = inp['pos_atoms'][:, :, 1] + torch.randn(N, L, 3, generator=gen)
    aa = torch.randint(-1, 24, (N, L), generator=gen)
    for rec in (~inp['context_mask'], torch.ones(N, L, dtype=torch.bool), torch.zeros(N, L, dtype=torch.bool)):
        args = (inp['pos_atoms'], R, t, aa, inp['chain_nb'], inp['res_nb'], inp['mask_atoms'], rec)
        want_pos, want_mask = OP.reconstruct_backbone_partially(*args, gold['bb_table'], gold['o_table'])
        got_pos, got_mask = P.reconstruct_backbone_partially(*[x.to(DEV) for x in args], bb_table=gold['bb_table'], o_table=gold['o_table'])
        assert torch.equal(got_mask.cpu(), want_mask)
        assert_atoms_close(got_pos.cpu(), want_pos)
        assert torch.equal(got_pos.cpu()[~rec], inp['pos_atoms'][~rec])                          # context atoms are copied bit for bit


def test_rmsd_and_rank_against_reference_fixture(gold):
    S = gold['structures']
    torch.testing.assert_close(ab_opt_b200.calc_per_rmsd(S.to(DEV)).cpu(), gold['rmsd'], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(ab_opt_b200.calc_avg_rmsd(S.to(DEV)).cpu(), torch.tensor(gold['avg_rmsd']), rtol=1e-5, atol=0)
    assert torch.equal(ab_opt_b200.rank_commoness(S.to(DEV), 5).cpu(), gold['rank'])
    assert torch.equal(ab_opt_b200.rank_commoness(S, 5), gold['rank'])                           # CPU in -> CPU out


@pytest.mark.parametrize('B,M', [(2, 1), (50, 33), (1000, 48)])
def test_rank_against_oracle(B, M):
    """`-n 1000` candidates (design_for_testset.py:471): indices must equal the oracle's wherever the oracle's own ordering is
    decided by more than fp32 summation noise; the scores agree to 1e-5 relative."""
    gen = torch.Generator().manual_seed(B)
    S = torch.randn(1, M, 3, generator=gen) * 10 + torch.randn(B, M, 3, generator=gen) * torch.rand(B, 1, 1, generator=gen) * 3
    k = min(B, 10)
    want_score = OP.commonness(S.double())
    order = torch.argsort(want_score)
    got = ab_opt_b200.rank_commoness(S.to(DEV), k).cpu()
    top = order[:k]
    ranked = want_score[order][:k + 1]
    if ((ranked[1:] - ranked[:-1]).abs() > 1e-5 * want_score.max()).all():
        assert torch.equal(got, top)
    else:           # near ties (B = 2: the two scores are the same number): what was picked must score like the oracle's picks
        torch.testing.assert_close(want_score[got], want_score[top], rtol=2e-5, atol=0)
    assert len(set(got.tolist())) == k
    torch.testing.assert_close(ab_opt_b200.calc_per_rmsd(S.to(DEV)).cpu().double(), OP.pairwise_rmsd(S.double()), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(ab_opt_b200.calc_avg_rmsd(S.to(DEV)).cpu().double(), OP.average_rmsd(S.double()), rtol=1e-5, atol=0)


def test_errors():
    S = torch.randn(4, 5, 3, device=DEV)
    with pytest.raises(ab_opt_b200.AboptError):
        ab_opt_b200.rank_commoness(S, 9)                                                         # k > B
    with pytest.raises(ab_opt_b200.AboptError):
        ab_opt_b200.calc_per_rmsd(S[:1])                                                         # a single structure: B - 1 = 0
