"""GPU parity of the pair featurisation: ab_opt_b200.PairEmbedding (one sm_100a kernel behind
abopt_pair_embed_forward) against the CPU oracle and the reference fixture tests/golden/pair_embed.npz.

Tolerance: fp32 against an fp64 evaluation of the oracle, err(cuda) <= 2 err(oracle fp32) + 5e-6, on every pair with i != j.
On the diagonal the inter-residue dihedral of the reference is +-acos(0.999999) = +-1.4e-3 rad with a sign decided by the
rounding of a triple product that is zero in exact arithmetic (geometry.py:268 with p0 == p3), so i == j gets 2e-3 absolute."""
import os

import numpy as np
import pytest
import torch

import ab_opt_b200
from oracle import pair_embed as PE

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def build(W, A):
    m = ab_opt_b200.PairEmbedding(64, A)
    m.load_state_dict(W, strict=True)
    return m.to(DEV).eval()


def run(mod, inp, sm=None, qm=None):
    c = {k: v.to(DEV) for k, v in inp.items()}
    return mod(c['aa'], c['res_nb'], c['chain_nb'], c['pos_atoms'], c['mask_atoms'],
               None if sm is None else sm.to(DEV), None if qm is None else qm.to(DEV))


def oracle(W, inp, sm, qm, dtype):
    Wd = {k: v.to(dtype) for k, v in W.items()}
    return PE.pair_embedding(Wd, inp['aa'], inp['res_nb'], inp['chain_nb'], inp['pos_atoms'].to(dtype), inp['mask_atoms'], sm, qm)


def check(got, W, inp, sm, qm):
    got = got.cpu()
    L = got.shape[1]
    z32, z64 = oracle(W, inp, sm, qm, torch.float32), oracle(W, inp, sm, qm, torch.float64)
    off = ~torch.eye(L, dtype=torch.bool)[None, :, :, None]
    e_got = ((got.double() - z64) * off).abs().max().item()
    e_ref = ((z32.double() - z64) * off).abs().max().item()
    assert e_got <= 2 * e_ref + 5e-6, f'off-diagonal: cuda err {e_got:.3e}, oracle fp32 err {e_ref:.3e}'
    assert (got.double() - z64).abs().max().item() <= 2e-3
    has_ca = inp['mask_atoms'][:, :, 1]
    assert (got[~has_ca] == 0).all() and (got.transpose(1, 2)[~has_ca] == 0).all()          # pair.py:99


@pytest.mark.parametrize('A', [15, 5])
@pytest.mark.parametrize('masked', [False, True])
def test_against_reference_fixture(golden_dir, A, masked):
    d = np.load(os.path.join(golden_dir, 'pair_embed.npz'))
    g = {k: torch.from_numpy(d[k]) if d[k].ndim else d[k].item() for k in d.files}
    W = PE.make_state_dict(g['seed_w'], A)
    inp = {k: g[k] for k in ('aa', 'res_nb', 'chain_nb', 'pos_atoms', 'mask_atoms')}
    m = g['context_mask'] if masked else None
    got = run(build(W, A), inp, m, m).cpu()
    ref = g[f'z_a{A}_' + ('masked' if masked else 'plain')]
    off = ~torch.eye(g['L'], dtype=torch.bool)[None, :, :, None]
    torch.testing.assert_close(got * off, ref * off, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(got, ref, rtol=0, atol=2e-3)
    check(got, W, inp, m, m)


@pytest.mark.parametrize('A,L', [(15, 100), (4, 65), (15, 1), (5, 64)])
def test_against_oracle_ragged(A, L):
    """Several 64-pair tiles with a ragged tail, three chains with numbering gaps, missing atoms, a residue without CA,
    padding rows; structure mask only (sequence kept), as `remove_structure` alone does (models/diffab.py:52-53)."""
    W = PE.make_state_dict(5, A)
    inp = PE.synthetic_complex(31 + L, 3, L)
    mod = build(W, A)
    check(run(mod, inp, inp['context_mask'], None), W, inp, inp['context_mask'], None)
    check(run(mod, inp, None, inp['context_mask']), W, inp, None, inp['context_mask'])


def test_empty_and_errors():
    W = PE.make_state_dict(5, 15)
    mod = build(W, 15)
    inp = PE.synthetic_complex(1, 2, 8)
    empty = {k: v[:0] for k, v in inp.items()}
    assert run(mod, empty).shape == (0, 8, 8, 64)
    with pytest.raises(ab_opt_b200.AboptError):                          # fewer atoms than the tables were built for
        short = dict(inp, pos_atoms=inp['pos_atoms'][:, :, :5], mask_atoms=inp['mask_atoms'][:, :, :5])
        run(mod, short)
    with pytest.raises(ab_opt_b200.AboptError):                          # no CPU fallback
        ab_opt_b200.PairEmbedding(64, 15)(inp['aa'], inp['res_nb'], inp['chain_nb'], inp['pos_atoms'], inp['mask_atoms'])


def test_full_size_properties():
    """C2 shapes (B=64, L=256, 15 atoms): determinism, independence of the complexes (a complex computed alone gives the same
    bits), masking, and the first / last complex against the oracle."""
    B, L, A = 64, 256, 15
    W = PE.make_state_dict(3, A)
    inp = PE.synthetic_complex(77, B, L)
    mod = build(W, A)
    m = inp['context_mask']
    z = run(mod, inp, m, m)
    assert torch.equal(z, run(mod, inp, m, m))
    for n in (0, 63):
        one = {k: v[n:n + 1] for k, v in inp.items()}
        z1 = run(mod, one, m[n:n + 1], m[n:n + 1])
        assert torch.equal(z1[0], z[n])
        check(z1, W, one, m[n:n + 1], m[n:n + 1])
    has_ca = inp['mask_atoms'][:, :, 1].to(DEV)
    assert (z[~has_ca] == 0).all()
    assert torch.isfinite(z).all()


# ------------------------------------------------------------------------------------------ ResidueEmbedding
def build_res(W, A):
    m = ab_opt_b200.ResidueEmbedding(128, A)
    m.load_state_dict(W, strict=True)
    return m.to(DEV).eval()


def run_res(mod, inp, ft, sm=None, qm=None):
    c = {k: v.to(DEV) for k, v in inp.items()}
    return mod(c['aa'], c['res_nb'], c['chain_nb'], c['pos_atoms'], c['mask_atoms'], ft.to(DEV),
               None if sm is None else sm.to(DEV), None if qm is None else qm.to(DEV))


def check_res(got, W, inp, ft, sm, qm):
    def orc(dtype):
        Wd = {k: v.to(dtype) for k, v in W.items()}
        return PE.residue_embedding(Wd, inp['aa'], inp['res_nb'], inp['chain_nb'], inp['pos_atoms'].to(dtype), inp['mask_atoms'], ft, sm, qm)
    x32, x64 = orc(torch.float32), orc(torch.float64)
    e_got = (got.cpu().double() - x64).abs().max().item()
    e_ref = (x32.double() - x64).abs().max().item()
    assert e_got <= 2 * e_ref + 2e-6, f'cuda err {e_got:.3e}, oracle fp32 err {e_ref:.3e}'
    assert (got.cpu()[~inp['mask_atoms'][:, :, 1]] == 0).all()                                   # residue.py:93


@pytest.mark.parametrize('A', [15, 5])
@pytest.mark.parametrize('masked', [False, True])
def test_residue_against_reference_fixture(golden_dir, A, masked):
    d = np.load(os.path.join(golden_dir, 'pair_embed.npz'))
    g = {k: torch.from_numpy(d[k]) if d[k].ndim else d[k].item() for k in d.files}
    W = PE.make_residue_state_dict(g['seed_w'] + 1, A)
    inp = {k: g[k] for k in ('aa', 'res_nb', 'chain_nb', 'pos_atoms', 'mask_atoms')}
    m = g['context_mask'] if masked else None
    got = run_res(build_res(W, A), inp, g['fragment_type'], m, m)
    torch.testing.assert_close(got.cpu(), g[f'x_a{A}_' + ('masked' if masked else 'plain')], rtol=1e-4, atol=2e-6)
    check_res(got, W, inp, g['fragment_type'], m, m)


@pytest.mark.parametrize('A,N,L', [(15, 3, 101), (4, 2, 7), (15, 1, 1), (5, 64, 256)])
def test_residue_against_oracle(A, N, L):
    """Row counts that are not multiples of the four-residue CTA pass, chain breaks and numbering gaps (termini), a residue
    without CA, padding, structure-only and sequence-only masks; the last case is the C2 batch."""
    W = PE.make_residue_state_dict(6, A)
    inp = PE.synthetic_complex(40 + L, N, L)
    ft = torch.randint(0, 10, (N, L), generator=torch.Generator().manual_seed(L))
    mod = build_res(W, A)
    m = inp['context_mask']
    check_res(run_res(mod, inp, ft, m, None), W, inp, ft, m, None)
    check_res(run_res(mod, inp, ft, None, m), W, inp, ft, None, m)
    a, b = run_res(mod, inp, ft, m, m), run_res(mod, inp, ft, m, m)
    assert torch.equal(a, b)
