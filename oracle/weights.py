"""Oracle: seeded synthetic FullDPM state-dicts with the reference's exact keys and shapes
(test infrastructure only).  Keys/shapes follow SURVEY.md section 8b "Weights"
(/root/reference/AbDock/src/modules/diffusion/dpm_full.py:37-68,134-147, encoders/ga.py:41-79).

numpy's legacy RandomState is used so that the same seed gives the same weights on every
machine (the golden fixtures store only inputs/outputs, not the 14 MB of weights).
"""
import math

import numpy as np
import torch

from .transitions import diffusion_buffers

F_DIM, C_DIM, H, D, P = 128, 64, 12, 32, 8
_BUFFER_CACHE = {}


def _lin(rs, W, name, n_out, n_in, bias=True):
    b = 1.0 / math.sqrt(n_in)
    W[name + '.weight'] = torch.from_numpy(rs.uniform(-b, b, (n_out, n_in)).astype(np.float32))
    if bias:
        W[name + '.bias'] = torch.from_numpy(rs.uniform(-b, b, (n_out,)).astype(np.float32))


def _ln(rs, W, name, n):
    W[name + '.gamma'] = torch.from_numpy((1 + 0.1 * rs.standard_normal(n)).astype(np.float32))
    W[name + '.beta'] = torch.from_numpy((0.1 * rs.standard_normal(n)).astype(np.float32))


def make_state_dict(seed=0, num_layers=6, flavour='abdock', num_bins=40, num_steps=100,
                    point_scale=1.0):
    """flavour 'abdock' adds the pRMSD head (+ prmsd.tobin.offset); 'abdesign' omits it."""
    rs = np.random.RandomState(seed)
    W = {}
    W['eps_net.current_sequence_embedding.weight'] = torch.from_numpy(
        rs.standard_normal((25, F_DIM)).astype(np.float32))
    _lin(rs, W, 'eps_net.res_feat_mixer.0', F_DIM, 2 * F_DIM)
    _lin(rs, W, 'eps_net.res_feat_mixer.2', F_DIM, F_DIM)
    for l in range(num_layers):
        p = f'eps_net.encoder.blocks.{l}.'
        W[p + 'spatial_coef'] = torch.from_numpy(
            (math.log(math.e - 1) + 0.3 * rs.standard_normal((1, 1, 1, H))).astype(np.float32))
        _lin(rs, W, p + 'proj_query', H * D, F_DIM, bias=False)
        _lin(rs, W, p + 'proj_key', H * D, F_DIM, bias=False)
        _lin(rs, W, p + 'proj_value', H * D, F_DIM, bias=False)
        _lin(rs, W, p + 'proj_pair_bias', H, C_DIM, bias=False)
        for nm in ('proj_query_point', 'proj_key_point', 'proj_value_point'):
            _lin(rs, W, p + nm, H * P * 3, F_DIM, bias=False)
            W[p + nm + '.weight'] *= point_scale
        _lin(rs, W, p + 'out_transform', F_DIM, H * C_DIM + H * D + H * P * 7)
        _ln(rs, W, p + 'layer_norm_1', F_DIM)
        for i in (0, 2, 4):
            _lin(rs, W, p + f'mlp_transition.{i}', F_DIM, F_DIM)
        _ln(rs, W, p + 'layer_norm_2', F_DIM)
    for head, n_out in (('eps_crd_net', 3), ('eps_rot_net', 3), ('eps_seq_net', 20)):
        _lin(rs, W, f'eps_net.{head}.0', F_DIM, F_DIM + 3)
        _lin(rs, W, f'eps_net.{head}.2', F_DIM, F_DIM)
        _lin(rs, W, f'eps_net.{head}.4', n_out, F_DIM)
    if flavour == 'abdock':
        _ln(rs, W, 'eps_net.prmsd_predictor.layer_norm', F_DIM + 3)
        _lin(rs, W, 'eps_net.prmsd_predictor.linear_1', F_DIM, F_DIM + 3)
        _lin(rs, W, 'eps_net.prmsd_predictor.linear_2', F_DIM, F_DIM)
        _lin(rs, W, 'eps_net.prmsd_predictor.linear_3', num_bins, F_DIM)
        W['prmsd.tobin.offset'] = torch.linspace(0.5, 19.5, num_bins)
    elif flavour != 'abdesign':
        raise ValueError(flavour)
    if num_steps not in _BUFFER_CACHE:
        _BUFFER_CACHE[num_steps] = diffusion_buffers(num_steps)
    W.update({k: v.clone() for k, v in _BUFFER_CACHE[num_steps].items()})
    W['_dummy'] = torch.empty(0)
    W['trans_rot._dummy'] = torch.empty(0)
    return W


def cast(W, dtype):
    """Floating tensors -> dtype (bool / long buffers untouched)."""
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in W.items()}


def synthetic_inputs(seed, N, L, gen_slices=((8, 12),), ragged=False, dtype=torch.float):
    """Seeded synthetic batch following SURVEY.md section 8d: res_feat, pair_feat ~ N(0,1);
    v = log of uniform rotations; p ~ N(0, 10 A); s ~ U{0..19}; mask_generate = given slices;
    mask_res all True, or (ragged) lengths U{0.75L..L} with padded aa = 21."""
    from .geometry import uniform_so3_from_gauss4
    rs = np.random.RandomState(seed)
    f = lambda *shape: torch.from_numpy(rs.standard_normal(shape).astype(np.float32))
    res_feat, pair_feat = f(N, L, F_DIM), f(N, L, L, C_DIM)
    v = uniform_so3_from_gauss4(f(N, L, 4))
    p = f(N, L, 3) * 10.0
    s = torch.from_numpy(rs.randint(0, 20, (N, L)).astype(np.int64))
    mask_generate = torch.zeros(N, L, dtype=torch.bool)
    for a, b in gen_slices:
        mask_generate[:, a:b] = True
    mask_res = torch.ones(N, L, dtype=torch.bool)
    if ragged:
        lens = rs.randint(int(0.75 * L), L + 1, (N,))
        for n, ln in enumerate(lens):
            mask_res[n, ln:] = False
            s[n, ln:] = 21
        mask_generate &= mask_res
    out = dict(v=v, p=p, s=s, res_feat=res_feat, pair_feat=pair_feat,
               mask_generate=mask_generate, mask_res=mask_res)
    return {k: (t.to(dtype) if t.is_floating_point() else t) for k, t in out.items()}
