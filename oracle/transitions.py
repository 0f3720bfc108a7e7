"""Oracle: variance schedule, IGSO(3) angle tables and the three reverse transitions
(test infrastructure only).

Restates /root/reference/AbDock/src/modules/diffusion/transition.py and the sampling half of
/root/reference/AbDock/src/modules/common/so3.py.  All randomness is an explicit INPUT
(`noise` tensors) so that the oracle, the reference (through a replayed torch.Generator) and
the CUDA path can be fed exactly the same draws; `draw_*` produce them in the reference's
ATen call order ("RNG draw order").
"""
import math

import torch
import torch.nn.functional as F

from .geometry import so3_exp, so3_log

NUM_BINS = 8192      # so3.py:73 default
NUM_AA = 20          # transition.py:165 default


# --------------------------------------------------------------------------- schedule
def variance_schedule(num_steps=100, s=0.01):
    """Cosine schedule buffers, each (T+1,) fp32.  transition.py:10-34."""
    T = num_steps
    t = torch.arange(0, T + 1, dtype=torch.float)
    f = torch.cos((math.pi / 2) * ((t / T) + s) / (1 + s)) ** 2
    alpha_bars = f / f[0]
    betas = torch.cat([torch.zeros(1), 1 - alpha_bars[1:] / alpha_bars[:-1]]).clamp_max(0.999)
    sig = torch.zeros_like(betas)
    for i in range(1, T + 1):
        sig[i] = ((1 - alpha_bars[i - 1]) / (1 - alpha_bars[i])) * betas[i]
    return {
        'betas': betas, 'alpha_bars': alpha_bars, 'alphas': 1 - betas, 'sigmas': torch.sqrt(sig),
        'sqrt_recip_alphas_cumprod': torch.sqrt(1. / alpha_bars),
        'sqrt_recipm1_alphas_cumprod': torch.sqrt(1. / alpha_bars - 1),
    }


def angular_tables(stddevs, std_threshold=0.1, num_bins=NUM_BINS, num_iters=1024):
    """Histogram approximation of the IGSO(3) angle density per sigma.  so3.py:73-109.

    Returns dict(stddevs (S,), approx_flag (S,) bool, X (S,bins), Y (S,bins)).
    """
    sd = torch.tensor(list(stddevs), dtype=torch.float)
    x = torch.linspace(0, math.pi, num_bins)
    l = torch.arange(0, num_iters)[None, :]
    c = ((1 - torch.cos(x)) / math.pi)[:, None]
    ratio = (torch.sin((l + 0.5) * x[:, None]) + 1e-6) / (torch.sin(x[:, None] / 2) + 1e-6)
    Y = []
    for e in sd.tolist():
        a = (2 * l + 1) * torch.exp(-l * (l + 1) * (e ** 2))
        Y.append(torch.nan_to_num((c * a * ratio).sum(1)).clamp_min(0))
    return {'stddevs': sd, 'approx_flag': sd <= std_threshold,
            'X': x[None].repeat(len(Y), 1), 'Y': torch.stack(Y)}


def diffusion_buffers(num_steps=100):
    """All non-learned FullDPM buffers under their state-dict keys.  dpm_full.py:135-147,
    transition.py:36-40,103-116,162-166."""
    sched = variance_schedule(num_steps)
    W = {}
    for mod in ('trans_rot', 'trans_pos', 'trans_seq'):
        for k, v in sched.items():
            W[f'{mod}.var_sched.{k}'] = v.clone()
    fwd = angular_tables(torch.sqrt(1 - sched['alpha_bars']).tolist())
    inv = angular_tables(sched['sigmas'].tolist())
    for name, tab in (('fwd', fwd), ('inv', inv)):
        for k, v in tab.items():
            W[f'trans_rot.angular_distrib_{name}.{k}'] = v
    W['position_mean'] = torch.zeros(1, 1, 3)
    W['position_scale'] = torch.full((1, 1, 1), 10.0)
    return W


# --------------------------------------------------------------------------- noise draws
def multinomial_from_exp(prob, expo):
    """torch.multinomial(prob, 1) given its internal Exp(1) draw: argmax(prob / q).
    (ATen native/Sampling: q = empty_like(p).exponential_(1); div; argmax.)"""
    return torch.argmax(prob / expo, dim=-1)


def draw_init_noise(N, L, gen, device='cpu', dtype=torch.float):
    """Draws of FullDPM.sample's initialisation, in order.  dpm_full.py:255-267."""
    g4 = torch.randn(N, L, 4, generator=gen, device=device, dtype=dtype)       # random_uniform_so3
    gp = torch.randn(N, L, 3, generator=gen, device=device, dtype=dtype)       # randn_like(p)
    s = torch.randint(0, 19, (N, L), generator=gen, device=device)             # randint_like(s, 0, 19)
    return {'g4': g4, 'gp': gp, 's_rand': s}


def draw_step_noise(N, L, gen, device='cpu', dtype=torch.float):
    """Draws of one reverse step, in order: so3.py:143 (u), so3.py:123 (multinomial ->
    exponential_ over (N*L, 8191)), so3.py:126 (rand_like), so3.py:131 (randn_like),
    transition.py:95 (randn_like p), transition.py:179 (multinomial -> exponential_ (N*L,20))."""
    M = N * L
    return {
        'u': torch.randn(N, L, 3, generator=gen, device=device, dtype=dtype),
        'expo_ang': torch.empty(M, NUM_BINS - 1, device=device, dtype=dtype).exponential_(1, generator=gen),
        'unif_ang': torch.rand(M, generator=gen, device=device, dtype=dtype),
        'gauss_ang': torch.randn(M, generator=gen, device=device, dtype=dtype),
        'z_pos': torch.randn(N, L, 3, generator=gen, device=device, dtype=dtype),
        'expo_seq': torch.empty(M, NUM_AA, device=device, dtype=dtype).exponential_(1, generator=gen),
    }


# --------------------------------------------------------------------------- rotation
def sample_angle(W, table, std_idx, expo, unif, gauss):
    """ApproxAngularDistribution.sample.  so3.py:111-138.  std_idx (M,) long."""
    p = f'trans_rot.angular_distrib_{table}.'
    X, Y, sd, flag = W[p + 'X'], W[p + 'Y'], W[p + 'stddevs'], W[p + 'approx_flag']
    prob = Y[std_idx].to(expo.dtype)
    b = multinomial_from_exp(prob[:, :-1], expo)
    Xd = X.to(expo.dtype)
    start = Xd[std_idx, b]
    width = Xd[std_idx, b + 1] - Xd[std_idx, b]
    hist = start + unif * width
    sdd = sd.to(expo.dtype)
    gau = (sdd[std_idx] * 2 + gauss * sdd[std_idx]).abs() % math.pi
    return torch.where(flag[std_idx], gau, hist)


def random_normal_so3(W, table, std_idx, noise):
    """so3.py:141-146.  std_idx (N,L)."""
    u = F.normalize(noise['u'], dim=-1)
    theta = sample_angle(W, table, std_idx.flatten(), noise['expo_ang'], noise['unif_ang'],
                         noise['gauss_ang']).reshape(std_idx.shape)
    return u * theta[..., None]


def rot_denoise(W, v_t, v_next, mask_generate, t, noise):
    """RotationTransition.denoise.  transition.py:146-160.  t (N,) long."""
    N, L = mask_generate.shape
    e = random_normal_so3(W, 'inv', t[:, None].expand(N, L), noise)
    e = torch.where((t > 1)[:, None, None].expand(N, L, 3), e, torch.zeros_like(e))
    v = so3_log(so3_exp(e) @ so3_exp(v_next))
    return torch.where(mask_generate[..., None].expand_as(v), v, v_t)


def rot_add_noise(W, v_0, mask_generate, t, noise):
    """RotationTransition.add_noise.  transition.py:120-144."""
    N, L = mask_generate.shape
    ab = W['trans_rot.var_sched.alpha_bars'].to(v_0.dtype)[t]
    c0 = torch.sqrt(ab).view(-1, 1, 1)
    e = random_normal_so3(W, 'fwd', t[:, None].expand(N, L), noise)
    v = so3_log(so3_exp(e) @ so3_exp(c0 * v_0))
    return torch.where(mask_generate[..., None].expand_as(v_0), v, v_0), e


# --------------------------------------------------------------------------- position
def pos_pred_noise_from_start(W, p_t, p_0, mask_generate, t):
    """transition.py:42-50 (used when obj == 'pred_x0', dpm_full.py:286-287)."""
    a = W['trans_pos.var_sched.sqrt_recip_alphas_cumprod'].to(p_t.dtype)[t].view(-1, 1, 1)
    b = W['trans_pos.var_sched.sqrt_recipm1_alphas_cumprod'].to(p_t.dtype)[t].view(-1, 1, 1)
    eps = (a * p_t - p_0) / b
    return torch.where(mask_generate[..., None].expand_as(p_t), eps, p_t)


def pos_denoise(W, p_t, eps_p, mask_generate, t, z):
    """PositionTransition.denoise.  transition.py:80-101."""
    dt = p_t.dtype
    al = W['trans_pos.var_sched.alphas'].to(dt)
    alpha = al[t].clamp_min(al[-2])
    alpha_bar = W['trans_pos.var_sched.alpha_bars'].to(dt)[t]
    sigma = W['trans_pos.var_sched.sigmas'].to(dt)[t].view(-1, 1, 1)
    c0 = (1.0 / torch.sqrt(alpha + 1e-8)).view(-1, 1, 1)
    c1 = ((1 - alpha) / torch.sqrt(1 - alpha_bar + 1e-8)).view(-1, 1, 1)
    z = torch.where((t > 1)[:, None, None].expand_as(p_t), z, torch.zeros_like(p_t))
    p = c0 * (p_t - c1 * eps_p) + sigma * z
    return torch.where(mask_generate[..., None].expand_as(p_t), p, p_t)


def pos_add_noise(W, p_0, mask_generate, t, e_rand):
    """PositionTransition.add_noise.  transition.py:62-78."""
    ab = W['trans_pos.var_sched.alpha_bars'].to(p_0.dtype)[t]
    p = torch.sqrt(ab).view(-1, 1, 1) * p_0 + torch.sqrt(1 - ab).view(-1, 1, 1) * e_rand
    return torch.where(mask_generate[..., None].expand_as(p_0), p, p_0)


# --------------------------------------------------------------------------- sequence
def one_hot_clamped(x, K, dtype):
    """layers.py:10-14: out-of-range classes (e.g. padding 21) give an all-zero row."""
    ok = (x >= 0) & (x < K)
    return (F.one_hot(x.clamp(0, K - 1), K) * ok[..., None]).to(dtype)


def seq_posterior(W, c_t, c_0, t):
    """transition.py:202-227; note alpha and alpha_bar are BOTH alpha_bars[t] (reference quirk)."""
    K = NUM_AA
    a = W['trans_seq.var_sched.alpha_bars'].to(c_t.dtype)[t][:, None, None]
    theta = ((a * c_t) + (1 - a) / K) * ((a * c_0) + (1 - a) / K)
    return theta / (theta.sum(-1, keepdim=True) + 1e-8)


def seq_denoise(W, s_t, c0_pred, mask_generate, t, expo):
    """AminoacidCategoricalTransition.denoise + _sample on ALL rows.  transition.py:170-181,229-245."""
    N, L = s_t.shape
    c_t = one_hot_clamped(s_t, NUM_AA, c0_pred.dtype)
    post = seq_posterior(W, c_t, c0_pred, t)
    post = torch.where(mask_generate[..., None].expand(post.shape), post, c_t)
    s_next = multinomial_from_exp(post.reshape(N * L, NUM_AA) + 1e-8, expo).view(N, L)
    return post, s_next


def seq_add_noise(W, s_0, mask_generate, t, expo):
    """transition.py:183-200."""
    N, L = s_0.shape
    dt = expo.dtype
    c_0 = one_hot_clamped(s_0, NUM_AA, dt)
    ab = W['trans_seq.var_sched.alpha_bars'].to(dt)[t][:, None, None]
    c = torch.where(mask_generate[..., None].expand(N, L, NUM_AA), ab * c_0 + (1 - ab) / NUM_AA, c_0)
    return c, multinomial_from_exp(c.reshape(N * L, NUM_AA) + 1e-8, expo).view(N, L)
