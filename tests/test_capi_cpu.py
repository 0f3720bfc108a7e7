"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol the header
declares, the Python mirror keeps the reference's state-dict layout, and the product refuses to run
without CUDA (no CPU fallback).  No compute is launched here."""
import os
import re

import pytest
import torch

import ab_opt_b200
from ab_opt_b200 import _capi
from oracle import weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module', autouse=True)
def built():
    from ab_opt_b200.build import build
    build()


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'abopt_b200.h')).read()
    declared = set(re.findall(r'\b(abopt_[a-z0-9_]+)\s*\(', header))
    declared -= {'abopt_model', 'abopt_config', 'abopt_step_noise', 'abopt_init_noise'}
    assert declared, 'no prototypes found in the header'
    lib = _capi.lib()
    for sym in sorted(declared):
        assert hasattr(lib, sym), f'{sym} declared in include/abopt_b200.h but not exported'
    assert declared == set(_capi.EXPORTS)
    assert lib.abopt_version() >= 100


@pytest.mark.parametrize('flavour', ['abdock', 'abdesign'])
def test_state_dict_is_reference_compatible(flavour):
    W = weights.make_state_dict(seed=0, num_layers=2, flavour=flavour)     # keys strict-load into the reference
    if flavour == 'abdock':
        m = ab_opt_b200.FullDPM(128, 64, 100, eps_net_opt=dict(num_layers=2), obj='pred_x0', num_bins=40)
    else:
        m = ab_opt_b200.FullDPMAbDesign(128, 64, 100, eps_net_opt=dict(num_layers=2))
    m.load_state_dict(W, strict=True)
    sd = m.state_dict()
    assert set(sd) == set(W)
    for k in W:
        assert sd[k].shape == W[k].shape and sd[k].dtype == W[k].dtype, k
    # the init-time buffers equal the reference's bit for bit (oracle buffers are pinned to it)
    fresh = ab_opt_b200.FullDPM(128, 64, 100, eps_net_opt=dict(num_layers=1), num_bins=40).state_dict()
    for k, v in fresh.items():
        if k.startswith('trans_') or k.startswith('position'):
            assert torch.equal(v, W[k]), k


def test_no_cpu_fallback():
    m = ab_opt_b200.FullDPM(128, 64, 100, eps_net_opt=dict(num_layers=1), num_bins=40)
    inp = weights.synthetic_inputs(0, 1, 8, gen_slices=((2, 5),))
    with pytest.raises(ab_opt_b200.AboptError):
        m.sample(inp['v'], inp['p'], inp['s'], inp['res_feat'], inp['pair_feat'], inp['mask_generate'], inp['mask_res'])
    enc = ab_opt_b200.GAEncoder(128, 64, 1)
    with pytest.raises(ab_opt_b200.AboptError):
        enc(torch.eye(3).expand(1, 8, 3, 3), inp['p'], inp['res_feat'], inp['pair_feat'], inp['mask_res'])


def test_unsupported_configuration_is_rejected():
    with pytest.raises(ValueError):
        ab_opt_b200.GABlock(128, 64, num_heads=8)
    with pytest.raises(ValueError):
        ab_opt_b200.EpsilonNet(256, 64, 2)


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the error path on a GPU-less host')
def test_model_create_fails_loudly_without_gpu():
    import ctypes
    cfg = _capi.Config(1, 100, 0, 0, 0.5, 19.5, 0, _capi.SCOPE_ENCODER)
    h = ctypes.c_void_p()
    rc = _capi.lib().abopt_model_create(ctypes.byref(cfg), 0, ctypes.byref(h))
    assert rc != 0 and _capi.lib().abopt_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the behaviour on a machine WITHOUT a CUDA device')
def test_callers_of_the_path_have_no_cpu_fallback():
    """The mirrors of the steps before / after the loop (PairEmbedding, ResidueEmbedding, reconstruct_backbone_partially,
    rank_commoness) accept CPU tensors like the reference's runners pass them, but only to move them to a GPU: without one they
    raise instead of computing anything on the host."""
    import ab_opt_b200
    aa = torch.zeros(1, 4, dtype=torch.long)
    pos, mask = torch.zeros(1, 4, 15, 3), torch.ones(1, 4, 15, dtype=torch.bool)
    with pytest.raises(ab_opt_b200.AboptError):
        ab_opt_b200.PairEmbedding(64, 15)(aa, aa, aa, pos, mask)
    with pytest.raises(ab_opt_b200.AboptError):
        ab_opt_b200.ResidueEmbedding(128, 15)(aa, aa, aa, pos, mask, aa)
    with pytest.raises(ab_opt_b200.AboptError):
        ab_opt_b200.rank_commoness(torch.zeros(3, 5, 3), 2)
    with pytest.raises(ab_opt_b200.AboptError):
        ab_opt_b200.reconstruct_backbone_partially(pos, torch.eye(3).expand(1, 4, 3, 3), torch.zeros(1, 4, 3), aa, aa, aa, mask,
                                                   torch.ones(1, 4, dtype=torch.bool), bb_table=torch.zeros(21, 3, 3), o_table=torch.zeros(21, 3))
    with pytest.raises(ValueError):                                     # the kernels are specialised for the reference configuration
        ab_opt_b200.PairEmbedding(32, 15)
    with pytest.raises(ValueError):
        ab_opt_b200.ResidueEmbedding(128, 16)
