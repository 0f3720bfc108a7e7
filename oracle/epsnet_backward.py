"""Oracle (TEST INFRASTRUCTURE ONLY): one whole training step -- FullDPM.forward and its backward -- with every gradient
written out by hand (no autograd): the losses' derivatives, the heads (pRMSD predictor included), the quaternion update, the
GAEncoder (oracle/ipa_backward.py per block) and the mixer / sequence embedding.  It is the blueprint of the CUDA backward
 and is checked against the gradients the unmodified reference computes with autograd
(tests/golden/train_backward.npz, AbDock flavour, both objectives) and against oracle.training.loss_and_grads.

Forward being differentiated: EpsilonNet.forward and FullDPM.forward of
  /root/reference/AbDock/src/modules/diffusion/dpm_full.py:70-112, 156-234   ('abdock': + pRMSD and, for obj pred_x0, distance loss)
  /root/reference/AbDesign/diffab/modules/diffusion/dpm_full.py:62-102, 138-190   ('abdesign': rot / pos / seq)
evaluated with the reference's grad-enabled log_rotation clamp (so3.py:12-17); the loss is the unweighted sum of the dict.
"""
import torch
import torch.nn.functional as F

from . import transitions as T
from .geometry import grad_enabled_semantics, quat_1ijk_to_rotation, so3_exp
from .ipa_backward import _layer_norm_backward, _linear_backward, ga_block_backward
from .ipa import ga_block, layer_norm
from .epsnet import has_prmsd, num_layers_of


def _mlp3_forward(W, p, x):
    a0 = F.linear(x, W[p + '0.weight'], W[p + '0.bias'])
    a1 = F.linear(F.relu(a0), W[p + '2.weight'], W[p + '2.bias'])
    return a0, a1, F.linear(F.relu(a1), W[p + '4.weight'], W[p + '4.bias'])


def _mlp3_backward(W, p, x, a0, a1, g_out, grads):
    g, grads[p + '4.weight'], grads[p + '4.bias'] = _linear_backward(g_out, F.relu(a1), W[p + '4.weight'])
    g, grads[p + '2.weight'], grads[p + '2.bias'] = _linear_backward(g * (a1 > 0), F.relu(a0), W[p + '2.weight'])
    g, grads[p + '0.weight'], grads[p + '0.bias'] = _linear_backward(g * (a0 > 0), x, W[p + '0.weight'])
    return g


def _quat_1ijk_backward(o, G):
    """U = rotation of the normalised quaternion (1, b, c, d) (geometry.py:215-233); G = d loss / d U -> d loss / d (b, c, d)."""
    n = torch.cat([torch.ones_like(o[..., :1]), o], -1)
    s = n.norm(dim=-1, keepdim=True)
    a, b, c, d = (n / s).unbind(-1)
    g = lambda i, j: G[..., i, j]
    tr = g(0, 0) + g(1, 1) + g(2, 2)
    ga = 2 * (a * tr + d * (g(1, 0) - g(0, 1)) + c * (g(0, 2) - g(2, 0)) + b * (g(2, 1) - g(1, 2)))
    gb = 2 * (b * (g(0, 0) - g(1, 1) - g(2, 2)) + c * (g(0, 1) + g(1, 0)) + d * (g(0, 2) + g(2, 0)) + a * (g(2, 1) - g(1, 2)))
    gc = 2 * (c * (g(1, 1) - g(0, 0) - g(2, 2)) + b * (g(0, 1) + g(1, 0)) + a * (g(0, 2) - g(2, 0)) + d * (g(1, 2) + g(2, 1)))
    gd = 2 * (d * (g(2, 2) - g(0, 0) - g(1, 1)) + a * (g(1, 0) - g(0, 1)) + b * (g(0, 2) + g(2, 0)) + c * (g(1, 2) + g(2, 1)))
    gu = torch.stack([ga, gb, gc, gd], -1)
    u = n / s
    gn = (gu - u * (u * gu).sum(-1, keepdim=True)) / s                   # u = n / |n|
    return gn[..., 1:]


def training_step_abdesign(W, v_0, p_0, s_0, res_feat, pair_feat, mask_generate, mask_res, t, noise):
    return training_step(W, v_0, p_0, s_0, res_feat, pair_feat, mask_generate, mask_res, t, noise, flavour='abdesign')


def training_step(W, v_0, p_0, s_0, res_feat, pair_feat, mask_generate, mask_res, t, noise, flavour='abdock', obj='pred_x0',
                  dist_min=0.5, dist_max=19.5):
    """-> (loss dict, {key: gradient} for every parameter, d / d res_feat, d / d pair_feat) of loss = sum of the loss dict."""
    N, L = mask_generate.shape
    dt = p_0.dtype
    grads = {}
    # ------------------------------------------------------------------ noising (constants of the step), dpm_full.py:146-160
    with grad_enabled_semantics():
        mean, scale = W['position_mean'].to(dt), W['position_scale'].to(dt)
        p_0n = (p_0 - mean) / scale
        R_0 = so3_exp(v_0)
        v_t, _ = T.rot_add_noise(W, v_0, mask_generate, t, noise)
        p_t = T.pos_add_noise(W, p_0n, mask_generate, t, noise['z_pos'])
        _, s_t = T.seq_add_noise(W, s_0, mask_generate, t, noise['expo_seq'])
    eps_p = noise['z_pos']
    beta = W['trans_pos.var_sched.betas'].to(dt)[t]
    # ------------------------------------------------------------------ EpsilonNet forward, keeping what the backward needs
    R = so3_exp(v_t)
    E = W['eps_net.current_sequence_embedding.weight']
    cat0 = torch.cat([res_feat, E[s_t]], -1)
    m_a = F.linear(cat0, W['eps_net.res_feat_mixer.0.weight'], W['eps_net.res_feat_mixer.0.bias'])
    xs = [F.linear(F.relu(m_a), W['eps_net.res_feat_mixer.2.weight'], W['eps_net.res_feat_mixer.2.bias'])]
    nl = num_layers_of(W)
    for l in range(nl):                                                   # only the block INPUTS are kept
        xs.append(ga_block(W, f'eps_net.encoder.blocks.{l}.', R, p_t, xs[-1], pair_feat, mask_res, materialize=False))
    t_embed = torch.stack([beta, torch.sin(beta), torch.cos(beta)], -1)[:, None, :].expand(N, L, 3)
    hcat = torch.cat([xs[-1], t_embed], -1)
    heads = {h: _mlp3_forward(W, f'eps_net.eps_{h}_net.', hcat) for h in ('crd', 'rot', 'seq')}
    gen = mask_generate[..., None].to(dt)
    eps_pos = torch.einsum('nlab,nlb->nla', R, heads['crd'][2]) * gen
    U = quat_1ijk_to_rotation(heads['rot'][2])
    R_pred = R @ U
    c_den = torch.softmax(heads['seq'][2], -1)
    # ------------------------------------------------------------------ losses (AbDesign dpm_full.py:162-188) and their derivatives
    mg = mask_generate.to(dt)
    denom = mg.sum() + 1e-8
    wgt = (mg / denom)[..., None]
    # rot: sum over columns of 1 - cos(col_pred, col_true) (cosine_embedding_loss, eps 1e-12 inside the square root)
    nx2, ny2 = R_pred.pow(2).sum(-2) + 1e-12, R_0.pow(2).sum(-2) + 1e-12  # per column
    cos = (R_pred * R_0).sum(-2) / torch.sqrt(nx2 * ny2)
    loss = {'rot': (((1 - cos).sum(-1)) * mg).sum() / denom}
    g_Rpred = -(R_0 / torch.sqrt(nx2 * ny2)[..., None, :] - R_pred * (cos / nx2)[..., None, :]) * wgt[..., None]
    # pos: |prediction - target|^2; the target is the noise (AbDesign :176), p_0 (AbDock pred_x0 :186-188) or p_noisy (AbDock
    # pred_noise, sic :189-191)
    pos_target = eps_p if flavour == 'abdesign' else (p_0n if obj == 'pred_x0' else p_t)
    loss['pos'] = ((eps_pos - pos_target).pow(2).sum(-1) * mg).sum() / denom
    g_eps_pos = 2 * (eps_pos - pos_target) * wgt
    g_hcat = torch.zeros_like(hcat)
    if flavour == 'abdock' and obj == 'pred_x0':
        # dist: SmoothL1 between the distance maps of p_pred and p_0 over rows of generated residues (dpm_full.py:369-378)
        dvec = eps_pos[:, :, None] - eps_pos[:, None]
        dp, dtrue = dvec.norm(dim=-1), torch.cdist(p_0n, p_0n)
        sel = (mask_generate[:, :, None] & mask_res[:, :, None] & mask_res[:, None, :]).to(dt)
        u = dp - dtrue
        loss['dist'] = (torch.where(u.abs() < 1, 0.5 * u * u, u.abs() - 0.5) * sel).sum() / sel.sum()
        g_dp = torch.where(u.abs() < 1, u, torch.sign(u)) * sel / sel.sum()
        g_dp = (g_dp + g_dp.transpose(1, 2)) / dp.clamp_min(1e-30) * (dp > 0)                      # d |p_i - p_j| = (p_i - p_j) / |.|
        g_eps_pos = g_eps_pos + (g_dp[..., None] * dvec).sum(2)
    if flavour == 'abdock' and has_prmsd(W):
        # pRMSD: cross entropy of the per-complex logits against the bin of the achieved RMSD (prmsd.py:53-69); the bin is an
        # argmin, so nothing flows back through the RMSD itself
        from .training import calc_rmsd
        pp = 'eps_net.prmsd_predictor.'
        ln = layer_norm(hcat, W[pp + 'layer_norm.gamma'], W[pp + 'layer_norm.beta'])
        b0 = F.linear(ln, W[pp + 'linear_1.weight'], W[pp + 'linear_1.bias'])
        b1 = F.linear(F.relu(b0), W[pp + 'linear_2.weight'], W[pp + 'linear_2.bias'])
        logits = F.linear(F.relu(b1), W[pp + 'linear_3.weight'], W[pp + 'linear_3.bias']).mean(dim=1)
        if obj == 'pred_x0':
            pred_p0 = eps_pos
        else:
            c0 = W['trans_pos.var_sched.sqrt_recip_alphas_cumprod'].to(dt)[t].view(-1, 1, 1)
            c1 = W['trans_pos.var_sched.sqrt_recipm1_alphas_cumprod'].to(dt)[t].view(-1, 1, 1)
            pred_p0 = torch.where(mask_generate[..., None].expand_as(p_0n), c0 * p_0n - c1 * eps_pos, p_0n)
        rmsd = calc_rmsd(pred_p0 * scale + mean, p_0n * scale + mean, mask_generate)
        offset = torch.linspace(dist_min, dist_max, logits.shape[-1]).to(dt)
        onehot = torch.zeros_like(logits).scatter_(-1, torch.argmin(torch.abs(rmsd.unsqueeze(-1) - offset), dim=-1, keepdim=True), 1.0)
        m0 = mask_generate[:, 0].to(dt)
        loss['prmsd'] = (-(onehot * F.log_softmax(logits, -1)).sum(-1) * m0).sum() / (m0.sum() + 1e-10)
        g_logits = (torch.softmax(logits, -1) - onehot) * (m0 / (m0.sum() + 1e-10))[:, None]
        g = (g_logits / L)[:, None, :].expand(N, L, -1)                                            # mean over ALL L rows
        g, grads[pp + 'linear_3.weight'], grads[pp + 'linear_3.bias'] = _linear_backward(g, F.relu(b1), W[pp + 'linear_3.weight'])
        g, grads[pp + 'linear_2.weight'], grads[pp + 'linear_2.bias'] = _linear_backward(g * (b1 > 0), F.relu(b0), W[pp + 'linear_2.weight'])
        g, grads[pp + 'linear_1.weight'], grads[pp + 'linear_1.bias'] = _linear_backward(g * (b0 > 0), ln, W[pp + 'linear_1.weight'])
        g, grads[pp + 'layer_norm.gamma'], grads[pp + 'layer_norm.beta'] = _layer_norm_backward(g, hcat, W[pp + 'layer_norm.gamma'])
        g_hcat += g
    # seq: KL(posterior(s_t, s_0) || posterior(s_t, c_denoised)), transition.py:202-227
    c_t, c_0 = T.one_hot_clamped(s_t, T.NUM_AA, dt), T.one_hot_clamped(s_0, T.NUM_AA, dt)
    a = W['trans_seq.var_sched.alpha_bars'].to(dt)[t][:, None, None]
    b = (1 - a) / T.NUM_AA
    A_t = a * c_t + b
    post_true = T.seq_posterior(W, c_t, c_0, t)
    theta = A_t * (a * c_den + b)
    S = theta.sum(-1, keepdim=True) + 1e-8
    post_pred = theta / S
    kl = torch.xlogy(post_true, post_true) - post_true * torch.log(post_pred + 1e-8)
    loss['seq'] = (kl.sum(-1) * mg).sum() / denom
    g_post = -post_true / (post_pred + 1e-8) * wgt
    g_theta = g_post / S - (g_post * theta).sum(-1, keepdim=True) / S.pow(2)
    g_cden = g_theta * A_t * a
    # ------------------------------------------------------------------ heads backward (dpm_full.py:84-100)
    g_seq = c_den * (g_cden - (c_den * g_cden).sum(-1, keepdim=True))                              # softmax
    g_crd = torch.einsum('nlba,nlb->nla', R, g_eps_pos * gen)                                      # eps_pos = R o, masked
    g_rot = _quat_1ijk_backward(heads['rot'][2], torch.einsum('nlba,nlbc->nlac', R, g_Rpred))      # R_pred = R U
    for h, g in (('crd', g_crd), ('rot', g_rot), ('seq', g_seq)):
        g_hcat += _mlp3_backward(W, f'eps_net.eps_{h}_net.', hcat, heads[h][0], heads[h][1], g, grads)
    # ------------------------------------------------------------------ encoder backward: one recompute-based block at a time
    g_x = g_hcat[..., :xs[-1].shape[-1]]
    g_pair = torch.zeros_like(pair_feat)
    for l in reversed(range(nl)):
        g_x, g_z, gw = ga_block_backward(W, f'eps_net.encoder.blocks.{l}.', R, p_t, xs[l], pair_feat, mask_res, g_x)
        g_pair += g_z
        grads.update(gw)
    # ------------------------------------------------------------------ mixer and sequence embedding backward (dpm_full.py:76-79)
    g, grads['eps_net.res_feat_mixer.2.weight'], grads['eps_net.res_feat_mixer.2.bias'] = _linear_backward(g_x, F.relu(m_a), W['eps_net.res_feat_mixer.2.weight'])
    g, grads['eps_net.res_feat_mixer.0.weight'], grads['eps_net.res_feat_mixer.0.bias'] = _linear_backward(g * (m_a > 0), cat0, W['eps_net.res_feat_mixer.0.weight'])
    Fd = res_feat.shape[-1]
    grads['eps_net.current_sequence_embedding.weight'] = torch.zeros_like(E).index_add_(0, s_t.reshape(-1), g[..., Fd:].reshape(-1, E.shape[1]))
    return loss, grads, g[..., :Fd], g_pair
