"""Oracle: EpsilonNet forward (test infrastructure only).

Restates EpsilonNet.forward of /root/reference/AbDock/src/modules/diffusion/dpm_full.py:70-112
(flavour 'abdock': pRMSD head, 5 outputs) and of
/root/reference/AbDesign/diffab/modules/diffusion/dpm_full.py:62-102 (flavour 'abdesign':
no pRMSD head, 4 outputs).  `W` is the FullDPM state-dict.
"""
import torch
import torch.nn.functional as F

from .geometry import so3_exp, so3_log, rotate_vector, quat_1ijk_to_rotation
from .ipa import ga_encoder, layer_norm


def _mlp3(W, p, x):
    h = F.relu(F.linear(x, W[p + '0.weight'], W[p + '0.bias']))
    h = F.relu(F.linear(h, W[p + '2.weight'], W[p + '2.bias']))
    return F.linear(h, W[p + '4.weight'], W[p + '4.bias'])


def num_layers_of(W):
    n = 0
    while f'eps_net.encoder.blocks.{n}.spatial_coef' in W:
        n += 1
    return n


def has_prmsd(W):
    return 'eps_net.prmsd_predictor.linear_1.weight' in W


def eps_net(W, v_t, p_t, s_t, res_feat, pair_feat, beta, mask_generate, mask_res,
            materialize=True, return_hidden=False):
    """Returns (v_next, R_next, eps_pos, c_denoised[, prmsd_logits])."""
    N, L = mask_res.shape
    R = so3_exp(v_t)                                                                  # :86
    emb = F.embedding(s_t, W['eps_net.current_sequence_embedding.weight'])
    x = torch.cat([res_feat, emb], -1)                                                # :89
    x = F.linear(F.relu(F.linear(x, W['eps_net.res_feat_mixer.0.weight'], W['eps_net.res_feat_mixer.0.bias'])),
                 W['eps_net.res_feat_mixer.2.weight'], W['eps_net.res_feat_mixer.2.bias'])
    x = ga_encoder(W, 'eps_net.encoder.', R, p_t, x, pair_feat, mask_res, num_layers_of(W), materialize)  # :90

    t_embed = torch.stack([beta, torch.sin(beta), torch.cos(beta)], -1)[:, None, :].expand(N, L, 3)  # :92
    h = torch.cat([x, t_embed], -1)

    gen3 = mask_generate[:, :, None].expand(N, L, 3)
    eps_pos = rotate_vector(R, _mlp3(W, 'eps_net.eps_crd_net.', h))                   # :96-97
    eps_pos = torch.where(gen3, eps_pos, torch.zeros_like(eps_pos))                   # :98

    U = quat_1ijk_to_rotation(_mlp3(W, 'eps_net.eps_rot_net.', h))                    # :101-102
    R_next = R @ U
    v_next = torch.where(gen3, so3_log(R_next), v_t)                                  # :104-105

    c_denoised = torch.softmax(_mlp3(W, 'eps_net.eps_seq_net.', h), -1)               # :108
    out = [v_next, R_next, eps_pos, c_denoised]
    if has_prmsd(W):                                                                  # :109-110, nn.py:180-188
        pp = 'eps_net.prmsd_predictor.'
        g = layer_norm(h, W[pp + 'layer_norm.gamma'], W[pp + 'layer_norm.beta'])
        g = F.relu(F.linear(g, W[pp + 'linear_1.weight'], W[pp + 'linear_1.bias']))
        g = F.relu(F.linear(g, W[pp + 'linear_2.weight'], W[pp + 'linear_2.bias']))
        g = F.linear(g, W[pp + 'linear_3.weight'], W[pp + 'linear_3.bias'])
        out.append(g.mean(dim=1))       # averaged over ALL L rows, padding included
    if return_hidden:
        out.append(x)
    return tuple(out)


def prmsd_score(logits, dist_min=0.5, dist_max=19.5):
    """pRMSDCa.compute_prmsd.  common/prmsd.py:31-47."""
    bounds = torch.linspace(dist_min, dist_max, logits.shape[-1], dtype=logits.dtype, device=logits.device)
    return (torch.softmax(logits, -1) * bounds).sum(-1)


def perplexity(post, mask_generate):
    """calc_perplexity (mean max-prob of softmax(post) over generated residues).  dpm_full.py:380-399."""
    mx = torch.softmax(post, -1).max(-1)[0] * mask_generate.to(post.dtype)
    return mx.sum(-1) / mask_generate.to(post.dtype).sum(-1)
