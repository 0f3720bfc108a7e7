"""EpsilonNet / FullDPM with the reference's constructor signatures, state-dict keys and call
signatures (/root/reference/AbDock/src/modules/diffusion/dpm_full.py:35-367; the AbDesign flavour
-- no pRMSD head, 3-tuples in the trajectory -- is selected with flavour='abdesign' or by the
alias classes at the bottom), executing on the sm_100a kernels of libabopt_b200.

RNG modes of sample()/optimize():
  rng='philox' (default) in-kernel counter-based Philox4x32-10; one seed is drawn from torch's
               default generator per call, so torch.manual_seed() still makes runs reproducible.
               Distributionally identical to the reference, not draw-for-draw.
  rng='torch'            every draw is made with the reference's own ATen calls in the reference's
               order on the same device and handed to the kernels ("parity" mode): a run seeded
               like the reference consumes the generator identically and reproduces it step by step.
"""
import ctypes
import os

import torch
import torch.nn as nn

from .. import _native
from ... import _capi
from ..common.layers import LayerNorm
from ..encoders.ga import GAEncoder
from .transition import RotationTransition, PositionTransition, AminoacidCategoricalTransition


class PerResiduePredictor(nn.Module):
    """Parameter container of the pRMSD head (common/nn.py:164-188)."""

    def __init__(self, no_bins, c_in, c_hidden):
        super().__init__()
        self.no_bins, self.c_in, self.c_hidden = no_bins, c_in, c_hidden
        self.layer_norm = LayerNorm(c_in)
        self.linear_1 = nn.Linear(c_in, c_hidden)
        self.linear_2 = nn.Linear(c_hidden, c_hidden)
        self.linear_3 = nn.Linear(c_hidden, no_bins)


class DistanceToBins(nn.Module):
    """Only the `offset` buffer (state-dict key prmsd.tobin.offset, common/layers.py:17-31)."""

    def __init__(self, dist_min, dist_max, num_bins):
        super().__init__()
        self.register_buffer('offset', torch.linspace(dist_min, dist_max, num_bins))


class pRMSDCa(nn.Module):
    def __init__(self, num_bins=20, dist_min=0.5, dist_max=19.5):
        super().__init__()
        self.num_bins, self.dist_min, self.dist_max = num_bins, dist_min, dist_max
        self.tobin = DistanceToBins(dist_min, dist_max, num_bins)


class EpsilonNet(_native.NativeOwner, nn.Module):
    _native_scope = _capi.SCOPE_EPSNET

    def __init__(self, res_feat_dim, pair_feat_dim, num_layers, no_bins=None, encoder_opt={}):
        """no_bins=None builds the AbDesign flavour (no pRMSD head, dpm_full.py:35 there)."""
        super().__init__()
        if res_feat_dim != 128 or pair_feat_dim != 64:
            raise ValueError('the sm_100a kernels are specialised for res_feat_dim=128, pair_feat_dim=64')
        self.current_sequence_embedding = nn.Embedding(25, res_feat_dim)
        self.res_feat_mixer = nn.Sequential(nn.Linear(res_feat_dim * 2, res_feat_dim), nn.ReLU(),
                                            nn.Linear(res_feat_dim, res_feat_dim))
        self.encoder = GAEncoder(res_feat_dim, pair_feat_dim, num_layers, **encoder_opt)

        def head(n_out):
            return nn.Sequential(nn.Linear(res_feat_dim + 3, res_feat_dim), nn.ReLU(),
                                 nn.Linear(res_feat_dim, res_feat_dim), nn.ReLU(), nn.Linear(res_feat_dim, n_out))
        self.eps_crd_net = head(3)
        self.eps_rot_net = head(3)
        self.eps_seq_net = nn.Sequential(*list(head(20)), nn.Softmax(dim=-1))
        self.no_bins = no_bins
        if no_bins is not None:
            self.prmsd_predictor = PerResiduePredictor(no_bins, res_feat_dim + 3, res_feat_dim)

    def _native_config(self):
        return _capi.Config(len(self.encoder.blocks), 100, int(self.no_bins is not None), int(self.no_bins or 0),
                            0.5, 19.5, 0, _capi.SCOPE_EPSNET)

    def _native_state(self):
        return {'eps_net.' + k: v for k, v in self.state_dict(keep_vars=True).items()}

    def _owner_native(self):
        owner = self.__dict__.get('_abopt_owner')
        return owner().native() if owner is not None and owner() is not None else self.native()

    @torch.no_grad()
    def forward(self, v_t, p_t, s_t, res_feat, pair_feat, beta, mask_generate, mask_res):
        """dpm_full.py:70-112 -> (v_next, R_next, eps_pos, c_denoised[, prmsd_logits])."""
        nm = self._owner_native()
        N, L = mask_res.shape
        v_t, p_t = _capi.cuda_f32(v_t, 'v_t'), _capi.cuda_f32(p_t, 'p_t')
        s_t = _capi.cuda_i64(s_t, 's_t')
        res_feat, pair_feat = _capi.cuda_f32(res_feat, 'res_feat'), _capi.cuda_f32(pair_feat, 'pair_feat')
        beta = _capi.cuda_f32(beta, 'beta')
        mg, mr = _capi.cuda_mask(mask_generate, 'mask_generate'), _capi.cuda_mask(mask_res, 'mask_res')
        if res_feat.shape != (N, L, 128) or pair_feat.shape != (N, L, L, 64) or beta.shape != (N,):
            raise ValueError('bad input shapes')
        dev = v_t.device
        v_next = torch.empty(N, L, 3, device=dev)
        R_next = torch.empty(N, L, 3, 3, device=dev)
        eps_pos = torch.empty(N, L, 3, device=dev)
        c_den = torch.empty(N, L, 20, device=dev)
        prm = torch.empty(N, self.no_bins, device=dev) if self.no_bins is not None else None
        _capi.check(_capi.lib().abopt_eps_net_forward(
            nm.handle, N, L, _capi.ptr(v_t), _capi.ptr(p_t), _capi.ptr(s_t), _capi.ptr(res_feat), _capi.ptr(pair_feat),
            _capi.ptr(beta), _capi.ptr(mg), _capi.ptr(mr), _capi.ptr(v_next), _capi.ptr(R_next), _capi.ptr(eps_pos),
            _capi.ptr(c_den), _capi.ptr(prm), _capi.stream_ptr(dev)))
        if prm is not None:
            return v_next, R_next, eps_pos, c_den, prm
        return v_next, R_next, eps_pos, c_den


class FullDPM(_native.NativeOwner, nn.Module):
    _native_scope = _capi.SCOPE_FULL

    def __init__(self, res_feat_dim, pair_feat_dim, num_steps, eps_net_opt={}, trans_rot_opt={}, trans_pos_opt={},
                 trans_seq_opt={}, position_mean=[0.0, 0.0, 0.0], position_scale=[10.0], obj='pred_noise',
                 num_bins=20, dist_min=0.5, dist_max=19.5, flavour='abdock', rng=None):
        super().__init__()
        assert obj in ['pred_x0', 'pred_noise']
        assert flavour in ('abdock', 'abdesign')
        self.flavour = flavour
        self.eps_net = EpsilonNet(res_feat_dim, pair_feat_dim, **eps_net_opt,
                                  no_bins=num_bins if flavour == 'abdock' else None)
        self.num_steps = num_steps
        self.trans_rot = RotationTransition(num_steps, **trans_rot_opt)
        self.trans_pos = PositionTransition(num_steps, **trans_pos_opt)
        self.trans_seq = AminoacidCategoricalTransition(num_steps, **trans_seq_opt)
        self.register_buffer('position_mean', torch.FloatTensor(position_mean).view(1, 1, -1))
        self.register_buffer('position_scale', torch.FloatTensor(position_scale).view(1, 1, -1))
        self.register_buffer('_dummy', torch.empty([0, ]))
        self.obj = obj
        self.num_bins, self.dist_min, self.dist_max = num_bins, dist_min, dist_max
        if flavour == 'abdock':
            self.prmsd = pRMSDCa(num_bins, dist_min=dist_min, dist_max=dist_max)
        self.rng = rng or os.environ.get('ABOPT_RNG', 'philox')
        self._rebind_children()

    def _rebind_children(self):
        _native.bind_owner(self, (self.eps_net, self.trans_rot, self.trans_pos, self.trans_seq))

    # ---------------------------------------------------------------- native handle
    def _native_config(self):
        abdock = self.flavour == 'abdock'
        return _capi.Config(len(self.eps_net.encoder.blocks), self.num_steps, int(abdock), self.num_bins if abdock else 0,
                            float(self.dist_min), float(self.dist_max), int(self.obj == 'pred_x0'), _capi.SCOPE_FULL)

    def _native_state(self):
        return dict(self.state_dict(keep_vars=True))

    def _normalize_position(self, p):
        return (p - self.position_mean) / self.position_scale

    def _unnormalize_position(self, p_norm):
        return p_norm * self.position_scale + self.position_mean

    def forward(self, v_0, p_0, s_0, res_feat, pair_feat, mask_generate, mask_res, denoise_structure, denoise_sequence,
                t=None, noise=None, rng=None):
        """dpm_full.py:156-234 (AbDesign :138-190).  Under torch.no_grad() (validate(), AbDock/train.py:141-149): the loss
        dict, values only.  With autograd enabled (train(), train.py:104-113): the same dict carrying a graph -- one
        autograd node whose backward is the hand-written sm_100a backward pass (abopt_loss_backward): `loss.backward()`
        fills .grad of every parameter and flows into res_feat / pair_feat."""
        needs_graph = torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters()) or res_feat.requires_grad
                                                   or pair_feat.requires_grad)
        if not needs_graph:
            with torch.no_grad():
                return self._loss_forward(v_0, p_0, s_0, res_feat, pair_feat, mask_generate, mask_res, denoise_structure,
                                          denoise_sequence, t, noise, rng)
        return self._train_forward(v_0, p_0, s_0, res_feat, pair_feat, mask_generate, mask_res, denoise_structure, denoise_sequence,
                                   t, noise, rng)

    def _loss_forward(self, v_0, p_0, s_0, res_feat, pair_feat, mask_generate, mask_res, denoise_structure, denoise_sequence,
                      t=None, noise=None, rng=None, _prepared=None):
        """dpm_full.py:156-234 (AbDesign :138-190): the loss dict of one training step, FORWARD ONLY -- the values are
        computed by the sm_100a kernels (abopt_loss_forward) and carry no autograd graph; the backward pass is not part
        of the native path yet.  `noise` (dict of the six draws of one step) replays given draws; otherwise rng='torch'
        makes the reference's ATen draws in the reference's order, rng='philox' draws inside the kernels."""
        nm = self.native()
        N, L = res_feat.shape[:2]
        dev = self._dummy.device
        if t is None:
            t = torch.randint(0, self.num_steps, (N,), dtype=torch.long, device=dev)
        t = _capi.cuda_i64(t, 't')
        v_0, p_0, s_0 = _capi.cuda_f32(v_0, 'v_0'), _capi.cuda_f32(p_0, 'p_0'), _capi.cuda_i64(s_0, 's_0')
        res_feat, pair_feat = _capi.cuda_f32(res_feat, 'res_feat'), _capi.cuda_f32(pair_feat, 'pair_feat')
        mg, mr = _capi.cuda_mask(mask_generate, 'mask_generate'), _capi.cuda_mask(mask_res, 'mask_res')
        if res_feat.shape != (N, L, 128) or pair_feat.shape != (N, L, L, 64) or t.shape != (N,):
            raise ValueError('bad input shapes')
        for nm_, x in (('v_0', v_0), ('p_0', p_0), ('s_0', s_0), ('res_feat', res_feat), ('pair_feat', pair_feat), ('t', t),
                       ('mask_generate', mg), ('mask_res', mr)):
            if x.device != dev:
                raise ValueError(f'{nm_} lives on {x.device}, the model on {dev}')
        # the reference indexes its (T+1)-entry schedules with t and raises IndexError outside [0, T]; a kernel would read out of bounds
        if N and (int(t.min()) < 0 or int(t.max()) > self.num_steps):
            raise IndexError(f't must be in [0, {self.num_steps}]')
        flags = (_capi.SAMPLE_STRUCTURE if denoise_structure else 0) | (_capi.SAMPLE_SEQUENCE if denoise_sequence else 0)
        rng = rng or self.rng
        M = N * L
        keep, nz, seed = None, None, 0
        if noise is None and rng == 'torch':      # transition.py:133 (so3.py:143,123,126,131), :74, :199
            noise = {}
            if denoise_structure:
                noise.update(u=torch.randn(N, L, 3, device=dev), expo_ang=torch.empty(M, 8191, device=dev).exponential_(1),
                             unif_ang=torch.rand(M, device=dev), gauss_ang=torch.randn(M, device=dev), z_pos=torch.randn(N, L, 3, device=dev))
            else:
                noise.update(u=torch.zeros(N, L, 3, device=dev), expo_ang=torch.ones(M, 8191, device=dev),
                             unif_ang=torch.zeros(M, device=dev), gauss_ang=torch.zeros(M, device=dev), z_pos=torch.zeros(N, L, 3, device=dev))
            noise['expo_seq'] = torch.empty(M, 20, device=dev).exponential_(1) if denoise_sequence else torch.ones(M, 20, device=dev)
        if noise is not None:
            keep = {k: _capi.cuda_f32(noise[k], k) for k in ('u', 'expo_ang', 'unif_ang', 'gauss_ang', 'z_pos', 'expo_seq')}
            nz = ctypes.byref(_capi.StepNoise(*[keep[k].data_ptr() for k in ('u', 'expo_ang', 'unif_ang', 'gauss_ang', 'z_pos', 'expo_seq')]))
        else:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        if _prepared is not None:          # the training path: hand the prepared call to the caller
            _prepared.update(nm=nm, N=N, L=L, v_0=v_0, p_0=p_0, s_0=s_0, res_feat=res_feat, pair_feat=pair_feat, mg=mg, mr=mr, flags=flags,
                             t=t, seed=seed, keep=keep, dev=dev)
            return None
        out = torch.zeros(5, device=dev)
        _capi.check(_capi.lib().abopt_loss_forward(
            nm.handle, N, L, _capi.ptr(v_0), _capi.ptr(p_0), _capi.ptr(s_0), _capi.ptr(res_feat), _capi.ptr(pair_feat),
            _capi.ptr(mg), _capi.ptr(mr), flags, _capi.ptr(t), seed, nz, _capi.ptr(out), _capi.stream_ptr(dev)))
        del keep
        return self._loss_dict(out)

    def _loss_dict(self, out):
        loss = {}
        if self.flavour == 'abdock':
            loss['prmsd'] = out[3]
            if self.obj == 'pred_x0':
                loss['dist'] = out[4]
        loss['rot'], loss['pos'], loss['seq'] = out[0], out[1], out[2]
        return loss

    # ---------------------------------------------------------------- training step
    def _noise_struct(self, prep):
        keep = prep['keep']
        if keep is None:
            return None
        return ctypes.byref(_capi.StepNoise(*[keep[k].data_ptr() for k in ('u', 'expo_ang', 'unif_ang', 'gauss_ang', 'z_pos', 'expo_seq')]))

    def _trainable(self):
        """[(FullDPM state-dict key, parameter)] of the parameters the backward pass produces gradients for."""
        return [(k, p) for k, p in self.named_parameters()]

    @torch.no_grad()
    def loss_and_grads(self, v_0, p_0, s_0, res_feat, pair_feat, mask_generate, mask_res, denoise_structure, denoise_sequence,
                       t=None, noise=None, rng=None, loss_weights=None):
        """One training step's forward AND backward in one call into the library (abopt_loss_backward): returns
        (loss dict, {state-dict key: gradient of sum_k w_k loss_k}, d / d res_feat, d / d pair_feat).  loss_weights: dict by
        loss name (configs/train/*.yml loss_weights), default all 1."""
        prep = {}
        self._loss_forward(v_0, p_0, s_0, res_feat, pair_feat, mask_generate, mask_res, denoise_structure, denoise_sequence, t, noise,
                           rng, _prepared=prep)
        return self._backward_call(prep, loss_weights)

    def _backward_call(self, prep, loss_weights=None, want_param_grads=True):
        nm, N, L, dev = prep['nm'], prep['N'], prep['L'], prep['dev']
        order = ('rot', 'pos', 'seq', 'prmsd', 'dist')
        lw = (ctypes.c_float * 5)(*[float((loss_weights or {}).get(k, 1.0)) for k in order])
        out = torch.zeros(5, device=dev)
        d_res, d_pair = torch.empty_like(prep['res_feat']), torch.empty_like(prep['pair_feat'])
        st = _capi.stream_ptr(dev)
        _capi.check(_capi.lib().abopt_loss_backward(
            nm.handle, N, L, _capi.ptr(prep['v_0']), _capi.ptr(prep['p_0']), _capi.ptr(prep['s_0']), _capi.ptr(prep['res_feat']),
            _capi.ptr(prep['pair_feat']), _capi.ptr(prep['mg']), _capi.ptr(prep['mr']), prep['flags'], _capi.ptr(prep['t']), prep['seed'],
            self._noise_struct(prep), lw, _capi.ptr(out), _capi.ptr(d_res), _capi.ptr(d_pair), st))
        grads = {}
        if want_param_grads:
            for k, p in self._trainable():
                g = torch.empty(p.numel(), device=dev)
                _capi.check(_capi.lib().abopt_model_get_grad(nm.handle, k.encode(), _capi.ptr(g), p.numel(), st))
                grads[k] = g.view(p.shape)
        return self._loss_dict(out), grads, d_res, d_pair

    def _train_forward(self, v_0, p_0, s_0, res_feat, pair_feat, mask_generate, mask_res, denoise_structure, denoise_sequence, t, noise, rng):
        prep = {}
        with torch.no_grad():
            self._loss_forward(v_0, p_0, s_0, res_feat, pair_feat, mask_generate, mask_res, denoise_structure, denoise_sequence, t, noise,
                               rng, _prepared=prep)
        prep['flags'] |= _capi.GRAD_SEMANTICS
        params = [p for _, p in self._trainable()]
        outs = _TrainStep.apply(self, prep, res_feat, pair_feat, *params)
        names = ('rot', 'pos', 'seq', 'prmsd', 'dist')
        have = self._loss_dict(torch.zeros(5))
        return {k: outs[i] for i, k in enumerate(names) if k in have}

    # ---------------------------------------------------------------- sampling
    @torch.no_grad()
    def sample(self, v, p, s, res_feat, pair_feat, mask_generate, mask_res, sample_structure=True, sample_sequence=True,
               pbar=False, **kwargs):
        """dpm_full.py:236-302.  Returns {t: [v, p_angstrom, s, prmsd, perplexity]} (AbDock flavour) or
        {t: (v, p_angstrom, s)} (AbDesign), t = num_steps..0; entries for t >= 1 live on the CPU, traj[0] on the device."""
        return self._run(v, p, s, 0, res_feat, pair_feat, mask_generate, mask_res, sample_structure, sample_sequence,
                         kwargs.get('rng', self.rng), kwargs.get('keep_trajectory', True), kwargs.get('seed'),
                         kwargs.get('batch_offset', 0), kwargs.get('batch_total'))

    @torch.no_grad()
    def optimize(self, v, p, s, opt_step: int, res_feat, pair_feat, mask_generate, mask_res, sample_structure=True,
                 sample_sequence=True, pbar=False, **kwargs):
        """dpm_full.py:304-367: noise to step `opt_step`, then denoise from there."""
        if not 1 <= int(opt_step) <= self.num_steps:
            raise ValueError('opt_step must be in [1, num_steps]')
        return self._run(v, p, s, int(opt_step), res_feat, pair_feat, mask_generate, mask_res, sample_structure,
                         sample_sequence, kwargs.get('rng', self.rng), kwargs.get('keep_trajectory', True), kwargs.get('seed'),
                         kwargs.get('batch_offset', 0), kwargs.get('batch_total'))

    def _run(self, v, p, s, opt_step, res_feat, pair_feat, mask_generate, mask_res, sample_structure, sample_sequence,
             rng, keep, seed=None, batch_offset=0, batch_total=None):
        """Extra keyword arguments of sample() / optimize() (batch sharding): `batch_offset` = index of v[0]
        in the global batch and `batch_total` = its size; `seed` = the Philox seed (default: drawn from torch's generator).
        With the same seed / generator state on every rank, the ranks' slices equal the unsharded run: the Philox counters
        are global rows, and rng='torch' draws the reference's full-batch tensors on every rank and uses its rows."""
        nm = self.native()
        L_ = _capi.lib()
        N, L = v.shape[:2]
        NT = int(batch_total) if batch_total is not None else N
        off = int(batch_offset)
        if off < 0 or off + N > NT:
            raise ValueError('batch_offset / batch_total do not contain this batch')
        _capi.check(L_.abopt_model_set_batch_offset(nm.handle, off))
        v, p = _capi.cuda_f32(v, 'v'), _capi.cuda_f32(p, 'p')
        s = _capi.cuda_i64(s, 's')
        res_feat, pair_feat = _capi.cuda_f32(res_feat, 'res_feat'), _capi.cuda_f32(pair_feat, 'pair_feat')
        mg, mr = _capi.cuda_mask(mask_generate, 'mask_generate'), _capi.cuda_mask(mask_res, 'mask_res')
        if res_feat.shape != (N, L, 128) or pair_feat.shape != (N, L, L, 64) or mg.shape != (N, L) or mr.shape != (N, L):
            raise ValueError('bad input shapes')
        dev = v.device
        abdock = self.flavour == 'abdock'
        optimize = opt_step > 0
        T0 = opt_step if optimize else self.num_steps
        flags = (_capi.SAMPLE_STRUCTURE if sample_structure else 0) | (_capi.SAMPLE_SEQUENCE if sample_sequence else 0)
        if T0 < 3:
            keep = True
        if keep:
            flags |= _capi.KEEP_TRAJECTORY
        tv = torch.empty(T0 + 1, N, L, 3, device=dev)
        tp = torch.empty(T0 + 1, N, L, 3, device=dev)
        ts = torch.empty(T0 + 1, N, L, dtype=torch.int64, device=dev)
        tpr = torch.zeros(T0 + 1, N, device=dev) if abdock else None
        tpl = torch.zeros(T0 + 1, N, device=dev) if abdock else None
        st = _capi.stream_ptr(dev)
        if rng == 'philox':
            seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if seed is None else int(seed)
            _capi.check(L_.abopt_sample_device(nm.handle, N, L, _capi.ptr(v), _capi.ptr(p), _capi.ptr(s), _capi.ptr(res_feat),
                                               _capi.ptr(pair_feat), _capi.ptr(mg), _capi.ptr(mr), flags, opt_step, seed, None,
                                               None, _capi.ptr(tv), _capi.ptr(tp), _capi.ptr(ts), _capi.ptr(tpr), _capi.ptr(tpl), st))
        elif rng == 'torch':
            flags |= _capi.KEEP_TRAJECTORY
            keep = True
            M = N * L
            MT = NT * L
            rows = slice(off * L, (off + N) * L)

            def mine(x, per_row=False):      # the full-batch draw of the reference -> the rows of this shard
                if NT == N:
                    return x
                return (x[rows] if per_row else x[off:off + N]).contiguous()

            def add_noise_draws():    # so3.py:143,123,126,131 ; transition.py:95 / :74 ; transition.py:179 / :199
                return dict(u=mine(torch.randn(NT, L, 3, device=dev)),
                            expo_ang=mine(torch.empty(MT, 8191, device=dev).exponential_(1), True),
                            unif_ang=mine(torch.rand(MT, device=dev), True), gauss_ang=mine(torch.randn(MT, device=dev), True),
                            z_pos=mine(torch.randn(NT, L, 3, device=dev)))

            def step_draws():
                d = add_noise_draws()
                d['expo_seq'] = mine(torch.empty(MT, 20, device=dev).exponential_(1), True)
                return d, _capi.StepNoise(*[d[k].data_ptr() for k in ('u', 'expo_ang', 'unif_ang', 'gauss_ang', 'z_pos', 'expo_seq')])
            if not optimize:          # dpm_full.py:255-267 (draws only happen for the enabled parts)
                g4 = mine(torch.randn(NT, L, 4, device=dev)) if sample_structure else torch.zeros(N, L, 4, device=dev)
                gp = mine(torch.randn(NT, L, 3, device=dev)) if sample_structure else torch.zeros_like(p)
                sr = (torch.randint_like(s, low=0, high=19) if NT == N else mine(torch.randint(0, 19, (NT, L), device=dev, dtype=s.dtype))) \
                    if sample_sequence else torch.zeros_like(s)
                init = _capi.InitNoise(g4.data_ptr(), gp.data_ptr(), sr.data_ptr(), None)
            else:                     # dpm_full.py:321-337
                d = {}
                if sample_structure:
                    d.update(add_noise_draws())
                else:
                    d.update(u=torch.zeros(N, L, 3, device=dev), expo_ang=torch.ones(M, 8191, device=dev),
                             unif_ang=torch.zeros(M, device=dev), gauss_ang=torch.zeros(M, device=dev), z_pos=torch.zeros_like(p))
                d['expo_seq'] = mine(torch.empty(MT, 20, device=dev).exponential_(1), True) if sample_sequence else torch.ones(M, 20, device=dev)
                add = _capi.StepNoise(*[d[k].data_ptr() for k in ('u', 'expo_ang', 'unif_ang', 'gauss_ang', 'z_pos', 'expo_seq')])
                init = _capi.InitNoise(None, None, None, ctypes.pointer(add))
            _capi.check(L_.abopt_sample_init(nm.handle, N, L, _capi.ptr(v), _capi.ptr(p), _capi.ptr(s), _capi.ptr(mg), flags, opt_step,
                                             0, ctypes.byref(init), _capi.ptr(tv[T0]), _capi.ptr(tp[T0]), _capi.ptr(ts[T0]), st))
            if abdock:
                tpl[T0] = 1.0
            for t in range(T0, 0, -1):
                d, nz = step_draws()
                _capi.check(L_.abopt_reverse_step(
                    nm.handle, N, L, t, int(optimize), _capi.ptr(tv[t]), _capi.ptr(tp[t]), _capi.ptr(ts[t]), _capi.ptr(res_feat),
                    _capi.ptr(pair_feat), _capi.ptr(mg), _capi.ptr(mr), flags, 0, ctypes.byref(nz), _capi.ptr(tv[t - 1]),
                    _capi.ptr(tp[t - 1]), _capi.ptr(ts[t - 1]), _capi.ptr(tpr[t - 1]) if abdock else None,
                    _capi.ptr(tpl[t - 1]) if abdock else None, st))
                del d
        else:
            raise ValueError("rng must be 'philox' or 'torch'")
        return self._pack_trajectory(tv, tp, ts, tpr, tpl, T0, keep, optimize)

    @torch.no_grad()
    def reverse_step(self, t, v_t, p_t, s_t, res_feat, pair_feat, mask_generate, mask_res, noise=None, seed=0,
                     sample_structure=True, sample_sequence=True, optimize=False):
        """One iteration of the sampling loop (dpm_full.py:274-298): traj[t] -> traj[t-1].  `p_t` in Angstrom.
        noise: dict with the six draws of a step (u, expo_ang, unif_ang, gauss_ang, z_pos, expo_seq) or None (Philox).
        Returns (v, p_angstrom, s[, prmsd, perplexity])."""
        nm = self.native()
        N, L = v_t.shape[:2]
        v_t, p_t, s_t = _capi.cuda_f32(v_t, 'v_t'), _capi.cuda_f32(p_t, 'p_t'), _capi.cuda_i64(s_t, 's_t')
        res_feat, pair_feat = _capi.cuda_f32(res_feat, 'res_feat'), _capi.cuda_f32(pair_feat, 'pair_feat')
        mg, mr = _capi.cuda_mask(mask_generate, 'mask_generate'), _capi.cuda_mask(mask_res, 'mask_res')
        dev = v_t.device
        abdock = self.flavour == 'abdock'
        flags = (_capi.SAMPLE_STRUCTURE if sample_structure else 0) | (_capi.SAMPLE_SEQUENCE if sample_sequence else 0)
        vo, po, so = torch.empty_like(v_t), torch.empty_like(p_t), torch.empty_like(s_t)
        pr = torch.empty(N, device=dev) if abdock else None
        pl = torch.empty(N, device=dev) if abdock else None
        nz = None
        if noise is not None:
            keep = {k: _capi.cuda_f32(noise[k], k) for k in ('u', 'expo_ang', 'unif_ang', 'gauss_ang', 'z_pos', 'expo_seq')}
            nz = ctypes.byref(_capi.StepNoise(*[keep[k].data_ptr() for k in ('u', 'expo_ang', 'unif_ang', 'gauss_ang', 'z_pos', 'expo_seq')]))
        _capi.check(_capi.lib().abopt_reverse_step(
            nm.handle, N, L, int(t), int(optimize), _capi.ptr(v_t), _capi.ptr(p_t), _capi.ptr(s_t), _capi.ptr(res_feat),
            _capi.ptr(pair_feat), _capi.ptr(mg), _capi.ptr(mr), flags, int(seed), nz, _capi.ptr(vo), _capi.ptr(po), _capi.ptr(so),
            _capi.ptr(pr), _capi.ptr(pl), _capi.stream_ptr(dev)))
        return (vo, po, so, pr, pl) if abdock else (vo, po, so)

    def _pack_trajectory(self, tv, tp, ts, tpr, tpl, T0, keep, optimize):
        """Device block -> the reference's dict.  One D2H copy of the whole block instead of the
        reference's per-step .cpu() synchronisations (dpm_full.py:299-300)."""
        abdock = self.flavour == 'abdock'
        slots = list(range(T0, 0, -1)) if keep else [T0]
        # fresh PINNED host tensors (torch's caching host allocator recycles the blocks of dropped trajectories), asynchronous
        # copies and one stream synchronisation: a pageable .cpu() of the 52 MB block goes through the driver's bounce buffer and
        # first-touch page faults (10-50 ms per sample, and the source of the run-to-run scatter of the bench line)
        def to_host(x):
            h = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
            h.copy_(x, non_blocking=True)
            return h
        hv, hp, hs = to_host(tv[1:]), to_host(tp[1:]), to_host(ts[1:])      # slots 1..T0
        hpr = to_host(tpr) if abdock else None
        hpl = to_host(tpl) if abdock else None
        torch.cuda.current_stream(tv.device).synchronize()
        traj = {}
        for t in slots:
            ent = [hv[t - 1], hp[t - 1], hs[t - 1]]
            if abdock:
                if t == T0:      # zeros_like / ones_like of the (N, L) sequence tensor (dpm_full.py:269,339)
                    ent += [torch.zeros_like(hs[t - 1]), torch.ones_like(hs[t - 1])]
                else:
                    ent += [hpr[t], hpl[t]]
            traj[t] = ent if (abdock and not optimize) else tuple(ent)
        last = [tv[0], tp[0], ts[0]]
        if abdock:   # sample(): prmsd / perplexity were already moved to the CPU (dpm_full.py:299); optimize(): stay on device
            last += [hpr[0], hpl[0]] if not optimize else [tpr[0], tpl[0]]
        traj[0] = last if (abdock and not optimize) else tuple(last)
        return traj


class _TrainStep(torch.autograd.Function):
    """FullDPM.forward as one autograd node.  forward: the loss dict, evaluated as the reference evaluates it with autograd
    enabled (abopt_loss_forward + ABOPT_GRAD_SEMANTICS).  backward: abopt_loss_backward with the incoming gradients of the five
    losses as loss weights (so any weighted sum of the dict differentiates correctly); it recomputes the forward from the saved
    inputs and replays the same noise (Philox seed, or the drawn tensors in parity mode)."""

    @staticmethod
    def forward(ctx, model, prep, res_feat, pair_feat, *params):
        out = torch.zeros(5, device=prep['dev'])
        _capi.check(_capi.lib().abopt_loss_forward(
            prep['nm'].handle, prep['N'], prep['L'], _capi.ptr(prep['v_0']), _capi.ptr(prep['p_0']), _capi.ptr(prep['s_0']),
            _capi.ptr(prep['res_feat']), _capi.ptr(prep['pair_feat']), _capi.ptr(prep['mg']), _capi.ptr(prep['mr']), prep['flags'],
            _capi.ptr(prep['t']), prep['seed'], model._noise_struct(prep), _capi.ptr(out), _capi.stream_ptr(prep['dev'])))
        ctx.model, ctx.prep = model, prep
        return tuple(out[i] for i in range(5))

    @staticmethod
    def backward(ctx, *g):
        model, prep = ctx.model, ctx.prep
        w = torch.stack([x if x is not None else torch.zeros((), device=prep['dev']) for x in g]).float().cpu().tolist()
        names = ('rot', 'pos', 'seq', 'prmsd', 'dist')
        _, grads, d_res, d_pair = model._backward_call(prep, {k: w[i] for i, k in enumerate(names)})
        pg = [grads[k] for k, _ in model._trainable()]
        return (None, None, d_res, d_pair, *pg)


class FullDPMAbDesign(FullDPM):
    """AbDesign flavour (AbDesign/diffab/modules/diffusion/dpm_full.py:106-130): no obj / pRMSD arguments."""

    def __init__(self, res_feat_dim, pair_feat_dim, num_steps, eps_net_opt={}, trans_rot_opt={}, trans_pos_opt={},
                 trans_seq_opt={}, position_mean=[0.0, 0.0, 0.0], position_scale=[10.0], rng=None):
        super().__init__(res_feat_dim, pair_feat_dim, num_steps, eps_net_opt, trans_rot_opt, trans_pos_opt, trans_seq_opt,
                         position_mean, position_scale, obj='pred_noise', flavour='abdesign', rng=rng)
