"""Oracle: FullDPM.forward, the loss dict of one training step (test infrastructure only).

Restates /root/reference/AbDock/src/modules/diffusion/dpm_full.py:156-234 (flavour 'abdock':
pRMSD loss, distance loss for obj='pred_x0', position loss on p_pred) and
/root/reference/AbDesign/diffab/modules/diffusion/dpm_full.py:138-190 (flavour 'abdesign':
rot / pos / seq, position loss on the predicted noise).  As everywhere in the oracle the random
draws are an explicit input: `noise` is one `transitions.draw_step_noise` record, which is exactly
the ATen draw order of the three add_noise calls (so3.py:143,123,126,131 ; transition.py:74 ;
transition.py:199 -> multinomial), and `t` is the (N,) step tensor.
"""
import torch
import torch.nn.functional as F

from . import transitions as T
from .epsnet import eps_net, has_prmsd
from .geometry import so3_exp


def rotation_matrix_cosine_loss(R_pred, R_true):
    """dpm_full.py:15-32: sum over the three columns of 1 - cos(column_pred, column_true)."""
    size = list(R_pred.shape[:-2])
    ncol = R_pred.numel() // 3
    RT_pred = R_pred.transpose(-2, -1).reshape(ncol, 3)
    RT_true = R_true.transpose(-2, -1).reshape(ncol, 3)
    ones = torch.ones([ncol], dtype=torch.long)
    loss = F.cosine_embedding_loss(RT_pred, RT_true, ones, reduction='none')
    return loss.reshape(size + [3]).sum(dim=-1)


def calc_rmsd(pred, target, mask):
    """pRMSDCa.calc_rmsd, common/prmsd.py:86-111."""
    m = mask.to(pred.dtype).unsqueeze(-1)
    sq = torch.sum((pred * m - target * m) ** 2, dim=-1)
    return torch.sqrt(torch.sum(sq, dim=-1) / torch.sum(mask, dim=-1))


def prmsd_loss(logits, rmsd, mask, dist_min=0.5, dist_max=19.5):
    """pRMSDCa.calc_prmsd_loss (common/prmsd.py:53-69) with the one-hot DistanceToBins (layers.py:48-51):
    cross entropy against the nearest of linspace(dist_min, dist_max, num_bins); mask is mask_generate[:, 0]."""
    offset = torch.linspace(dist_min, dist_max, logits.shape[-1]).to(logits.dtype)
    idx = torch.argmin(torch.abs(rmsd.unsqueeze(-1) - offset), dim=-1, keepdim=True)
    onehot = torch.zeros_like(logits).scatter_(-1, idx, 1.0)
    err = -torch.sum(onehot * F.log_softmax(logits, dim=-1), dim=-1)
    return (err * mask).sum() / (mask.sum() + 1e-10)


def calc_dist_loss(p_pred, p_true, mask_generate, mask_res):
    """dpm_full.py:369-378: SmoothL1 between the two distance maps over rows of generated residues."""
    dp, dt = torch.cdist(p_pred, p_pred), torch.cdist(p_true, p_true)
    mm = mask_res[:, :, None] & mask_res[:, None, :]
    sel = mask_generate[:, :, None].expand_as(dp) & mm
    return F.smooth_l1_loss(torch.masked_select(dp, sel), torch.masked_select(dt, sel), reduction='none').mean()


def loss_forward(W, v_0, p_0, s_0, res_feat, pair_feat, mask_generate, mask_res, denoise_structure, denoise_sequence,
                 t, noise, flavour='abdock', obj='pred_x0', dist_min=0.5, dist_max=19.5, materialize=True, taps=None):
    """FullDPM.forward.  p_0 in Angstrom.  Returns {'rot','pos','seq'[,'prmsd'][,'dist']} of 0-dim tensors."""
    N, L = mask_generate.shape
    mean, scale = W['position_mean'].to(p_0.dtype), W['position_scale'].to(p_0.dtype)
    p_0 = (p_0 - mean) / scale                                                           # :160
    R_0 = so3_exp(v_0)
    if denoise_structure:                                                                 # :162-167
        v_noisy, _ = T.rot_add_noise(W, v_0, mask_generate, t, noise)
        p_noisy = T.pos_add_noise(W, p_0, mask_generate, t, noise['z_pos'])
        eps_p = noise['z_pos']
    else:
        v_noisy, p_noisy, eps_p = v_0.clone(), p_0.clone(), torch.zeros_like(p_0)
    if denoise_sequence:                                                                  # :174-178
        _, s_noisy = T.seq_add_noise(W, s_0, mask_generate, t, noise['expo_seq'])
    else:
        s_noisy = s_0.clone()
    beta = W['trans_pos.var_sched.betas'].to(p_0.dtype)[t]
    out = eps_net(W, v_noisy, p_noisy, s_noisy, res_feat, pair_feat, beta, mask_generate, mask_res, materialize=materialize)
    R_pred, p_pred, c_den = out[1], out[2], out[3]
    if taps is not None:
        taps.update(v_noisy=v_noisy, p_noisy=p_noisy, s_noisy=s_noisy, R_pred=R_pred, p_pred=p_pred, c_denoised=c_den)
    mg = mask_generate.to(p_0.dtype)
    denom = mg.sum() + 1e-8
    loss = {}
    if flavour == 'abdock':
        if obj == 'pred_x0':                                                              # :186-191
            p_true, pred_p0 = p_0, p_pred
        else:                      # the reference compares the predicted noise with p_noisy here (sic)
            p_true = p_noisy
            # pred_start_from_noise(p_0, p_pred, ...) -- called with p_0 in the p_t slot, transition.py:52-60
            c0 = W['trans_pos.var_sched.sqrt_recip_alphas_cumprod'].to(p_0.dtype)[t].view(-1, 1, 1)
            c1 = W['trans_pos.var_sched.sqrt_recipm1_alphas_cumprod'].to(p_0.dtype)[t].view(-1, 1, 1)
            pred_p0 = torch.where(mask_generate[..., None].expand_as(p_0), c0 * p_0 - c1 * p_pred, p_0)
        if has_prmsd(W):
            rmsd = calc_rmsd(pred_p0 * scale + mean, p_0 * scale + mean, mask_generate)   # :196-197
            loss['prmsd'] = prmsd_loss(out[4], rmsd, mask_generate[:, 0].to(p_0.dtype), dist_min, dist_max)
        if obj == 'pred_x0':
            loss['dist'] = calc_dist_loss(p_pred, p_true, mask_generate, mask_res)        # :210-212
        pos_target = p_true
    else:
        pos_target = eps_p                                                                # AbDesign :176
    loss['rot'] = (rotation_matrix_cosine_loss(R_pred, R_0) * mg).sum() / denom
    loss['pos'] = (F.mse_loss(p_pred, pos_target, reduction='none').sum(dim=-1) * mg).sum() / denom
    post_true = T.seq_posterior(W, T.one_hot_clamped(s_noisy, T.NUM_AA, p_0.dtype), T.one_hot_clamped(s_0, T.NUM_AA, p_0.dtype), t)
    log_post_pred = torch.log(T.seq_posterior(W, T.one_hot_clamped(s_noisy, T.NUM_AA, p_0.dtype), c_den, t) + 1e-8)
    kl = F.kl_div(input=log_post_pred, target=post_true, reduction='none', log_target=False).sum(dim=-1)
    loss['seq'] = (kl * mg).sum() / denom
    return loss


def loss_and_grads(W, v_0, p_0, s_0, res_feat, pair_feat, mask_generate, mask_res, denoise_structure, denoise_sequence, t, noise,
                   loss_weights=None, **kw):
    """One training step's forward AND backward as the reference runs it with autograd enabled (train.py: loss =
    sum_k w_k loss_k; loss.backward()): `loss_forward` under `grad_enabled_semantics`, then torch autograd through the oracle.
    Returns (loss dict, {state-dict key: gradient} for every floating parameter that receives one, d loss / d res_feat,
    d loss / d pair_feat).  This is the target the CUDA backward will be held to."""
    from .geometry import grad_enabled_semantics
    buffers = ('trans_', 'position_', '_dummy', 'prmsd.tobin')
    Wg = {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() and not k.startswith(buffers) else v) for k, v in W.items()}
    rf, pf = res_feat.detach().clone().requires_grad_(True), pair_feat.detach().clone().requires_grad_(True)
    with torch.enable_grad(), grad_enabled_semantics():
        loss = loss_forward(Wg, v_0, p_0, s_0, rf, pf, mask_generate, mask_res, denoise_structure, denoise_sequence, t, noise, **kw)
        total = sum((loss_weights or {}).get(k, 1.0) * v for k, v in loss.items())
        total.backward()
    grads = {k: v.grad for k, v in Wg.items() if torch.is_tensor(v) and v.requires_grad and v.grad is not None}
    return {k: v.detach() for k, v in loss.items()}, grads, rf.grad, pf.grad
