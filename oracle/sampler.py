"""Oracle: FullDPM.sample / FullDPM.optimize reverse-diffusion loops (test infrastructure only).

Restates /root/reference/AbDock/src/modules/diffusion/dpm_full.py:236-367 (and the AbDesign
mirror :193-319).  Noise comes either from a torch.Generator (drawn in the reference's order)
or from a caller-supplied tape (list of dicts) for teacher-forced replays.
"""
import torch

from . import transitions as T
from .epsnet import eps_net, has_prmsd, prmsd_score, perplexity
from .geometry import uniform_so3_from_gauss4


def _norm(W, p):
    return (p - W['position_mean'].to(p.dtype)) / W['position_scale'].to(p.dtype)      # dpm_full.py:148-150


def _unnorm(W, p):
    return p * W['position_scale'].to(p.dtype) + W['position_mean'].to(p.dtype)        # dpm_full.py:152-154


def reverse_step(W, t, v_t, p_t, s_t, res_feat, pair_feat, mask_generate, mask_res, noise,
                 obj='pred_noise', materialize=True):
    """One iteration of the loop body on NORMALISED positions.  dpm_full.py:274-298.

    Returns dict with the network outputs and the next state (all tensors).
    """
    N, L = mask_res.shape
    dt = v_t.dtype
    beta = W['trans_pos.var_sched.betas'].to(dt)[t].expand(N)
    tt = torch.full((N,), t, dtype=torch.long, device=v_t.device)
    net = eps_net(W, v_t, p_t, s_t, res_feat, pair_feat, beta, mask_generate, mask_res, materialize)
    v_net, R_net, p_pred, c_den = net[:4]
    eps_p = T.pos_pred_noise_from_start(W, p_t, p_pred, mask_generate, tt) if obj == 'pred_x0' else p_pred
    v_next = T.rot_denoise(W, v_t, v_net, mask_generate, tt, noise)
    p_next = T.pos_denoise(W, p_t, eps_p, mask_generate, tt, noise['z_pos'])
    post, s_next = T.seq_denoise(W, s_t, c_den, mask_generate, tt, noise['expo_seq'])
    out = dict(v_net=v_net, R_net=R_net, p_pred=p_pred, c_denoised=c_den, eps_p=eps_p,
               v_next=v_next, p_next=p_next, s_next=s_next, post=post)
    if has_prmsd(W):
        out['prmsd'] = prmsd_score(net[4])
        out['prmsd_logits'] = net[4]
    return out


def sample(W, v, p, s, res_feat, pair_feat, mask_generate, mask_res, num_steps=100,
           sample_structure=True, sample_sequence=True, obj='pred_noise', gen=None, tape=None,
           materialize=True, start_step=None, record=None, stop_at=0):
    """FullDPM.sample (start_step None) or FullDPM.optimize (start_step = opt_step).

    `tape`: optional dict {'init': {...}, t: {...}} of pre-drawn noise; otherwise drawn from
    `gen` in the reference order.  Returns traj {t: [v, p_angstrom, s(, prmsd, perplexity)]}.
    `record`, if a dict, receives the per-step reverse_step outputs keyed by t.
    `stop_at`: stop once traj[stop_at] exists (0 = run to completion).
    """
    N, L = v.shape[:2]
    dt = v.dtype
    abdock = has_prmsd(W)
    p = _norm(W, p)
    gen3 = mask_generate[:, :, None].expand(N, L, 3)
    if start_step is None:                                                     # sample(): :254-267
        T0 = num_steps
        nz = tape['init'] if tape is not None else T.draw_init_noise(N, L, gen, dtype=dt)
        if sample_structure:
            v_init = torch.where(gen3, uniform_so3_from_gauss4(nz['g4']), v)
            p_init = torch.where(gen3, nz['gp'], p)
        else:
            v_init, p_init = v, p
        s_init = torch.where(mask_generate, nz['s_rand'], s) if sample_sequence else s
    else:                                                                       # optimize(): :321-337
        T0 = start_step
        tt = torch.full((N,), T0, dtype=torch.long, device=v.device)
        nz = tape['init'] if tape is not None else T.draw_step_noise(N, L, gen, dtype=dt)
        if sample_structure:
            v_noisy, _ = T.rot_add_noise(W, v, mask_generate, tt, nz)
            v_init = torch.where(gen3, v_noisy, v)
            p_init = torch.where(gen3, T.pos_add_noise(W, p, mask_generate, tt, nz['z_pos']), p)
        else:
            v_init, p_init = v, p
        if sample_sequence:
            _, s_noisy = T.seq_add_noise(W, s, mask_generate, tt, nz['expo_seq'])
            s_init = torch.where(mask_generate, s_noisy, s)
        else:
            s_init = s

    first = [v_init, _unnorm(W, p_init), s_init]
    if abdock:
        first += [torch.zeros_like(s_init), torch.ones_like(s_init)]
    traj = {T0: first}
    for t in range(T0, stop_at, -1):
        v_t, p_t, s_t = traj[t][:3]
        p_t = _norm(W, p_t)
        nz = tape[t] if tape is not None else T.draw_step_noise(N, L, gen, dtype=dt)
        # reference quirk: optimize() feeds the network's position output straight to denoise,
        # ignoring `obj` (dpm_full.py:351-356) -- only sample() converts x0 -> eps (:286-289).
        st = reverse_step(W, t, v_t, p_t, s_t, res_feat, pair_feat, mask_generate, mask_res, nz,
                          obj=obj if start_step is None else 'pred_noise', materialize=materialize)
        if record is not None:
            record[t] = st
        v_next, p_next, s_next = st['v_next'], st['p_next'], st['s_next']
        if not sample_structure:
            v_next, p_next = v_t, p_t
        if not sample_sequence:
            s_next = s_t
        nxt = [v_next, _unnorm(W, p_next), s_next]
        if abdock:
            gm = mask_generate if start_step is None else None   # optimize() passes no mask (:357)
            ppl = perplexity(st['post'], gm if gm is not None else torch.ones_like(mask_generate))
            nxt += [st['prmsd'], ppl]
        traj[t - 1] = nxt
    return traj
