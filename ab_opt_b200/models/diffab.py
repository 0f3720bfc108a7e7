"""DiffusionAntibodyDesign with the reference's constructor (cfg), state-dict keys (residue_embed.* / pair_embed.* /
diffusion.*) and call signatures (/root/reference/AbDock/src/models/diffab.py:19-171; AbDesign: diffab/models/diffab.py).

sample() / optimize() run as ONE device-resident call into libabopt_b200 (abopt_design_device): the batch's atoms go in,
ResidueEmbedding, PairEmbedding, the backbone frames and the whole reverse-diffusion loop run on the device and only the
trajectory comes back -- res_feat / pair_feat never exist as host-visible tensors.  encode() / forward() compose the module
mirrors the way the reference does."""
import torch
import torch.nn as nn

from .. import _capi
from ..modules.diffusion.dpm_full import FullDPM, FullDPMAbDesign
from ..modules.encoders.pair import PairEmbedding, ResidueEmbedding

resolution_to_num_atoms = {'backbone+CB': 5, 'full': 15}      # models/diffab.py:13-16 (max_num_heavyatoms = 15)
CA, C_, N_ = 1, 2, 0                                          # BBHeavyAtom (utils/protein/constants.py)


def _get(cfg, key, default=None):
    if hasattr(cfg, 'get'):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


class DiffusionAntibodyDesign(nn.Module):

    def __init__(self, cfg, flavour='abdock'):
        super().__init__()
        self.cfg = cfg
        num_atoms = resolution_to_num_atoms[_get(cfg, 'resolution', 'full')]
        self.residue_embed = ResidueEmbedding(_get(cfg, 'res_feat_dim'), num_atoms)
        self.pair_embed = PairEmbedding(_get(cfg, 'pair_feat_dim'), num_atoms)
        dopt = dict(_get(cfg, 'diffusion'))
        if flavour == 'abdock':          # models/diffab.py:30-37
            self.diffusion = FullDPM(_get(cfg, 'res_feat_dim'), _get(cfg, 'pair_feat_dim'), **dopt, num_bins=_get(cfg, 'num_bins', 20),
                                     dist_min=_get(cfg, 'dist_min', 0.5), dist_max=_get(cfg, 'dist_max', 19.5))
        else:                            # AbDesign: diffab/models/diffab.py:30-34
            self.diffusion = FullDPMAbDesign(_get(cfg, 'res_feat_dim'), _get(cfg, 'pair_feat_dim'), **dopt)

    # ---------------------------------------------------------------- the reference's composition (training / callers of encode)
    def encode(self, batch, remove_structure, remove_sequence):
        """models/diffab.py:39-86 -> res_feat (N,L,128), pair_feat (N,L,L,64), R (N,L,3,3), p (N,L,3)."""
        context_mask = torch.logical_and(batch['mask_heavyatom'][:, :, CA], ~batch['generate_flag'])
        structure_mask = context_mask if remove_structure else None
        sequence_mask = context_mask if remove_sequence else None
        a = (batch['aa'], batch['res_nb'], batch['chain_nb'], batch['pos_heavyatom'], batch['mask_heavyatom'])
        res_feat = self.residue_embed(*a, batch['fragment_type'], structure_mask=structure_mask, sequence_mask=sequence_mask)
        pair_feat = self.pair_embed(*a, structure_mask=structure_mask, sequence_mask=sequence_mask)
        pos = batch['pos_heavyatom']
        ca = pos[:, :, CA]
        e1 = pos[:, :, C_] - ca                                      # construct_3d_basis, geometry.py:47-69
        e1 = e1 / (torch.linalg.norm(e1, dim=-1, keepdim=True) + 1e-6)
        v2 = pos[:, :, N_] - ca
        u2 = v2 - (e1 * v2).sum(-1, keepdim=True) * e1
        e2 = u2 / (torch.linalg.norm(u2, dim=-1, keepdim=True) + 1e-6)
        R = torch.stack([e1, e2, torch.linalg.cross(e1, e2, dim=-1)], dim=-1)
        return res_feat, pair_feat, R, ca

    @staticmethod
    def _so3vec(R):                                                  # rotation_to_so3vec, so3.py:10-30,60-63 (no_grad clamp)
        tr = R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2]
        cos_t = ((tr - 1) / 2).clamp_min(-1.0)
        coef = (torch.acos(cos_t) + 1e-8) / (2 * torch.sqrt(1 - cos_t * cos_t) + 2e-8)
        A = coef[..., None, None] * (R - R.transpose(-1, -2))
        return torch.stack([A[..., 1, 2], A[..., 2, 0], A[..., 0, 1]], -1)

    def forward(self, batch):
        """models/diffab.py:88-113 (losses only; see FullDPM.forward)."""
        mask_generate = batch['generate_flag']
        if _get(self.cfg, 'mask_ratio_min', False):                  # models/diffab.py:97-100,166-180
            ratio = float(torch.empty(1).uniform_(_get(self.cfg, 'mask_ratio_min'), _get(self.cfg, 'mask_ratio_max')))
            rnd = torch.bernoulli(torch.zeros_like(mask_generate.float()).fill_(ratio)).bool()
            mask_generate = torch.logical_and(mask_generate, rnd)
            batch['generate_flag'] = mask_generate
        ts, tq = _get(self.cfg, 'train_structure', True), _get(self.cfg, 'train_sequence', True)
        res_feat, pair_feat, R_0, p_0 = self.encode(batch, remove_structure=ts, remove_sequence=tq)
        return self.diffusion(self._so3vec(R_0), p_0, batch['aa'], res_feat, pair_feat, mask_generate, batch['mask'],
                              denoise_structure=ts, denoise_sequence=tq)

    # ---------------------------------------------------------------- the fused device-resident path
    def _design(self, batch, opt_step, opt):
        dpm = self.diffusion
        nm, pe, re = dpm.native(), self.pair_embed.native(), self.residue_embed.native()
        aa = _capi.cuda_i64(batch['aa'], 'aa')
        N, L = aa.shape
        res_nb, chain_nb = _capi.cuda_i64(batch['res_nb'], 'res_nb'), _capi.cuda_i64(batch['chain_nb'], 'chain_nb')
        ft = _capi.cuda_i64(batch['fragment_type'], 'fragment_type')
        pos = _capi.cuda_f32(batch['pos_heavyatom'], 'pos_heavyatom')
        ma = _capi.cuda_mask(batch['mask_heavyatom'], 'mask_heavyatom')
        gen, mask = _capi.cuda_mask(batch['generate_flag'], 'generate_flag'), _capi.cuda_mask(batch['mask'], 'mask')
        A = pos.shape[2]
        if pos.shape != (N, L, A, 3) or ma.shape != (N, L, A) or gen.shape != (N, L) or mask.shape != (N, L):
            raise ValueError('bad input shapes')
        dev = aa.device
        abdock = dpm.flavour == 'abdock'
        optimize = opt_step > 0
        T0 = opt_step if optimize else dpm.num_steps
        flags = (_capi.SAMPLE_STRUCTURE if opt.get('sample_structure', True) else 0) | \
                (_capi.SAMPLE_SEQUENCE if opt.get('sample_sequence', True) else 0) | _capi.KEEP_TRAJECTORY
        tv = torch.empty(T0 + 1, N, L, 3, device=dev)
        tp = torch.empty(T0 + 1, N, L, 3, device=dev)
        ts = torch.empty(T0 + 1, N, L, dtype=torch.int64, device=dev)
        tpr = torch.zeros(T0 + 1, N, device=dev) if abdock else None
        tpl = torch.zeros(T0 + 1, N, device=dev) if abdock else None
        seed = opt.get('seed')
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if seed is None else int(seed)
        _capi.check(_capi.lib().abopt_model_set_batch_offset(nm.handle, int(opt.get('batch_offset', 0))))
        _capi.check(_capi.lib().abopt_design_device(
            nm.handle, pe.handle, re.handle, N, L, A, _capi.ptr(aa), _capi.ptr(res_nb), _capi.ptr(chain_nb), _capi.ptr(pos),
            _capi.ptr(ma), _capi.ptr(ft), _capi.ptr(gen), _capi.ptr(mask), flags, opt_step, seed, _capi.ptr(tv), _capi.ptr(tp),
            _capi.ptr(ts), _capi.ptr(tpr), _capi.ptr(tpl), _capi.stream_ptr(dev)))
        return dpm._pack_trajectory(tv, tp, ts, tpr, tpl, T0, True, optimize)

    @torch.no_grad()
    def sample(self, batch, sample_opt={'sample_structure': True, 'sample_sequence': True, 'contig': ''}):
        """models/diffab.py:115-141."""
        if sample_opt.get('sample_sequence', False) and sample_opt.get('contig', '') != '':      # generate_mask_from_str, :184-206
            start, end = (int(x) for x in sample_opt['contig'].split('-'))
            m = torch.zeros_like(batch['generate_flag'], dtype=torch.bool)
            m[..., start - 1:end] = True
            batch['generate_flag'] = torch.logical_and(batch['generate_flag'], m)
        return self._design(batch, 0, sample_opt)

    @torch.no_grad()
    def optimize(self, batch, opt_step, optimize_opt={'sample_structure': True, 'sample_sequence': True}):
        """models/diffab.py:143-171."""
        if not 1 <= int(opt_step) <= self.diffusion.num_steps:
            raise ValueError('opt_step must be in [1, num_steps]')
        return self._design(batch, int(opt_step), optimize_opt)
