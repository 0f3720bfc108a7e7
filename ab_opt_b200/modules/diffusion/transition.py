"""Variance schedule and the three transition modules with the reference's names, constructor
signatures and buffers (/root/reference/AbDock/src/modules/diffusion/transition.py).  `denoise` /
`pred_noise_from_start` run on the libabopt_b200 kernels; noise is drawn with the SAME ATen calls,
in the same order, as the reference, so a seeded run consumes the generator identically."""
import numpy as np
import torch
import torch.nn as nn

from ... import _capi
from .. import _native
from ..common.so3 import ApproxAngularDistribution


class VarianceSchedule(nn.Module):
    """Cosine schedule (transition.py:10-34)."""

    def __init__(self, num_steps=100, s=0.01):
        super().__init__()
        T = num_steps
        t = torch.arange(0, num_steps + 1, dtype=torch.float)
        f_t = torch.cos((np.pi / 2) * ((t / T) + s) / (1 + s)) ** 2
        alpha_bars = f_t / f_t[0]
        betas = torch.cat([torch.zeros([1]), 1 - (alpha_bars[1:] / alpha_bars[:-1])], dim=0).clamp_max(0.999)
        sigmas = torch.zeros_like(betas)
        for i in range(1, betas.size(0)):
            sigmas[i] = ((1 - alpha_bars[i - 1]) / (1 - alpha_bars[i])) * betas[i]
        self.register_buffer('betas', betas)
        self.register_buffer('alpha_bars', alpha_bars)
        self.register_buffer('alphas', 1 - betas)
        self.register_buffer('sigmas', torch.sqrt(sigmas))
        self.register_buffer('sqrt_recip_alphas_cumprod', torch.sqrt(1. / alpha_bars))
        self.register_buffer('sqrt_recipm1_alphas_cumprod', torch.sqrt(1. / alpha_bars - 1))


class _Transition(_native.Owned, nn.Module):
    """Transitions are leaves of FullDPM; the owner injects itself so they can reach its handle."""

    def _owner(self):
        owner = self.__dict__.get('_abopt_owner')
        if owner is None:
            raise _capi.AboptError('transition modules execute through their FullDPM (native handle owner); '
                                   'construct them via ab_opt_b200 FullDPM')
        return owner()


class PositionTransition(_Transition):

    def __init__(self, num_steps, var_sched_opt={}):
        super().__init__()
        self.var_sched = VarianceSchedule(num_steps, **var_sched_opt)

    @torch.no_grad()
    def pred_noise_from_start(self, p_t, p_0, mask_generate, t):
        """transition.py:42-50."""
        nm = self._owner().native()
        N, L = mask_generate.shape
        p_t, p_0 = _capi.cuda_f32(p_t, 'p_t'), _capi.cuda_f32(p_0, 'p_0')
        mg, tt = _capi.cuda_mask(mask_generate, 'mask_generate'), _capi.cuda_i64(t, 't')
        out = torch.empty_like(p_t)
        _capi.check(_capi.lib().abopt_pos_pred_noise_from_start(nm.handle, N, L, _capi.ptr(p_t), _capi.ptr(p_0), _capi.ptr(mg),
                                                               _capi.ptr(tt), _capi.ptr(out), _capi.stream_ptr(p_t.device)))
        return out

    @torch.no_grad()
    def denoise(self, p_t, eps_p, mask_generate, t):
        """transition.py:80-101; draws randn_like(p_t) like the reference."""
        nm = self._owner().native()
        N, L = mask_generate.shape
        p_t, eps_p = _capi.cuda_f32(p_t, 'p_t'), _capi.cuda_f32(eps_p, 'eps_p')
        mg, tt = _capi.cuda_mask(mask_generate, 'mask_generate'), _capi.cuda_i64(t, 't')
        z = torch.randn_like(p_t)
        out = torch.empty_like(p_t)
        _capi.check(_capi.lib().abopt_pos_denoise(nm.handle, N, L, _capi.ptr(p_t), _capi.ptr(eps_p), _capi.ptr(mg), _capi.ptr(tt),
                                                  _capi.ptr(z), _capi.ptr(out), _capi.stream_ptr(p_t.device)))
        return out


class RotationTransition(_Transition):

    def __init__(self, num_steps, var_sched_opt={}, angular_distrib_fwd_opt={}, angular_distrib_inv_opt={}):
        super().__init__()
        self.var_sched = VarianceSchedule(num_steps, **var_sched_opt)
        c1 = torch.sqrt(1 - self.var_sched.alpha_bars)
        self.angular_distrib_fwd = ApproxAngularDistribution(c1.tolist(), **angular_distrib_fwd_opt)
        self.angular_distrib_inv = ApproxAngularDistribution(self.var_sched.sigmas.tolist(), **angular_distrib_inv_opt)
        self.register_buffer('_dummy', torch.empty([0, ]))

    @torch.no_grad()
    def denoise(self, v_t, v_next, mask_generate, t):
        """transition.py:146-160.  Draw order = so3.py:143,123,126,131."""
        nm = self._owner().native()
        N, L = mask_generate.shape
        v_t, v_next = _capi.cuda_f32(v_t, 'v_t'), _capi.cuda_f32(v_next, 'v_next')
        mg, tt = _capi.cuda_mask(mask_generate, 'mask_generate'), _capi.cuda_i64(t, 't')
        dev = v_t.device
        u = torch.randn(N, L, 3, device=dev)
        expo = torch.empty(N * L, 8191, device=dev).exponential_(1)       # the draw inside torch.multinomial
        unif = torch.rand(N * L, device=dev)
        gauss = torch.randn(N * L, device=dev)
        out = torch.empty_like(v_t)
        _capi.check(_capi.lib().abopt_rot_denoise(nm.handle, N, L, _capi.ptr(v_t), _capi.ptr(v_next), _capi.ptr(mg), _capi.ptr(tt),
                                                  _capi.ptr(u), _capi.ptr(expo), _capi.ptr(unif), _capi.ptr(gauss),
                                                  _capi.ptr(out), _capi.stream_ptr(dev)))
        return out


class AminoacidCategoricalTransition(_Transition):

    def __init__(self, num_steps, num_classes=20, var_sched_opt={}):
        super().__init__()
        if num_classes != 20:
            raise ValueError('the CUDA kernels are specialised for 20 amino-acid classes')
        self.num_classes = num_classes
        self.var_sched = VarianceSchedule(num_steps, **var_sched_opt)

    @torch.no_grad()
    def denoise(self, x_t, c_0_pred, mask_generate, t):
        """transition.py:229-245 -> (post (N,L,20), x_next (N,L))."""
        nm = self._owner().native()
        N, L = mask_generate.shape
        x_t, c0 = _capi.cuda_i64(x_t, 'x_t'), _capi.cuda_f32(c_0_pred, 'c_0_pred')
        mg, tt = _capi.cuda_mask(mask_generate, 'mask_generate'), _capi.cuda_i64(t, 't')
        expo = torch.empty(N * L, 20, device=c0.device).exponential_(1)
        post = torch.empty_like(c0)
        x_next = torch.empty_like(x_t)
        _capi.check(_capi.lib().abopt_seq_denoise(nm.handle, N, L, _capi.ptr(x_t), _capi.ptr(c0), _capi.ptr(mg), _capi.ptr(tt),
                                                  _capi.ptr(expo), _capi.ptr(post), _capi.ptr(x_next),
                                                  _capi.stream_ptr(c0.device)))
        return post, x_next
