"""PairEmbedding with the reference's constructor signature, state-dict keys and call signature
(/root/reference/AbDock/src/modules/encoders/pair.py:10-101; AbDesign: diffab/modules/encoders/pair.py), executing as ONE
sm_100a kernel of libabopt_b200 (csrc/k_pair_embed.cu) through the C ABI.  No PyTorch arithmetic happens here."""
import ctypes as C

import torch
import torch.nn as nn

from ... import _capi
from .. import _native


class AngularEncoding(nn.Module):
    """Holds the `freq_bands` buffer of the reference (common/layers.py:85-95); the encoding runs inside the kernel."""

    def __init__(self, num_funcs=3):
        super().__init__()
        if num_funcs != 3:
            raise ValueError('the CUDA kernel implements the reference default num_funcs=3 only')
        self.num_funcs = num_funcs
        self.register_buffer('freq_bands', torch.FloatTensor([i + 1 for i in range(num_funcs)]
                                                             + [1. / (i + 1) for i in range(num_funcs)]))

    def get_out_dim(self, in_dim):
        return in_dim * (1 + 2 * 2 * self.num_funcs)


class _NativeEmbed:
    """Owns one abopt_pair_embed / abopt_res_embed handle (`kind` = 'pair' | 'res')."""

    def __init__(self, kind, max_num_atoms, device, tensors):
        L = _capi.lib()
        self.kind = kind
        create, set_tensor, finalize = (getattr(L, f'abopt_{kind}_embed_{f}') for f in ('create', 'set_tensor', 'finalize'))
        dev = torch.device(device)
        if dev.type != 'cuda':
            raise _capi.AboptError('ab_opt_b200 runs on CUDA devices only (no CPU fallback); got device ' + str(dev))
        self.index = dev.index if dev.index is not None else torch.cuda.current_device()
        h = C.c_void_p()
        _capi.check(create(max_num_atoms, self.index, C.byref(h)))
        self.handle = h
        try:
            for key, t in tensors.items():
                t = t.detach().float().contiguous()
                _capi.check(set_tensor(h, key.encode(), _capi.ptr(t), t.numel(), 1 if t.is_cuda else 0))
            _capi.check(finalize(h))
        except Exception:
            getattr(L, f'abopt_{kind}_embed_destroy')(h)
            self.handle = None
            raise

    def __del__(self):
        if getattr(self, 'handle', None) is not None and _capi._lib is not None:
            getattr(_capi._lib, f'abopt_{self.kind}_embed_destroy')(self.handle)
            self.handle = None


def _native_embed(module, kind):
    """The handle behind `module`, rebuilt whenever a parameter changes version, storage or device."""
    state = module.state_dict(keep_vars=True)
    fp = tuple((k, t.data_ptr(), t._version, str(t.device)) for k, t in state.items())
    cached = module.__dict__.get('_native_cache')
    if cached is not None and cached[0] == fp:
        return cached[1]
    devs = {t.device for t in state.values()}
    if len(devs) != 1:
        raise _capi.AboptError(f'parameters live on several devices: {devs}')
    nm = _NativeEmbed(kind, module.max_num_atoms, devs.pop(), state)
    module.__dict__['_native_cache'] = (fp, nm)
    return nm


def _common_inputs(aa, res_nb, chain_nb, pos_atoms, mask_atoms, structure_mask, sequence_mask):
    aa = _capi.cuda_i64(aa, 'aa'); res_nb = _capi.cuda_i64(res_nb, 'res_nb'); chain_nb = _capi.cuda_i64(chain_nb, 'chain_nb')
    pos = _capi.cuda_f32(pos_atoms, 'pos_atoms'); mask = _capi.cuda_mask(mask_atoms, 'mask_atoms')
    sm = _capi.cuda_mask(structure_mask, 'structure_mask') if structure_mask is not None else None
    qm = _capi.cuda_mask(sequence_mask, 'sequence_mask') if sequence_mask is not None else None
    N, L = aa.shape
    A = pos.shape[2] if pos.dim() == 4 else -1
    if pos.shape != (N, L, A, 3) or mask.shape != (N, L, A) or res_nb.shape != (N, L) or chain_nb.shape != (N, L) \
            or any(m is not None and m.shape != (N, L) for m in (sm, qm)):
        raise ValueError(f'bad shapes: aa {tuple(aa.shape)} pos_atoms {tuple(pos.shape)} mask_atoms {tuple(mask.shape)}')
    return aa, res_nb, chain_nb, pos, mask, sm, qm, N, L, A


class PairEmbedding(nn.Module):

    def __init__(self, feat_dim, max_num_atoms, max_aa_types=22, max_relpos=32):
        super().__init__()
        if (feat_dim, max_aa_types, max_relpos) != (64, 22, 32) or not 4 <= max_num_atoms <= 15:
            raise ValueError('the sm_100a kernel is specialised for the reference configuration: feat_dim 64, 22 amino-acid '
                             'types, max_relpos 32, 4..15 atoms per residue')
        self.max_num_atoms, self.max_aa_types, self.max_relpos = max_num_atoms, max_aa_types, max_relpos
        self.aa_pair_embed = nn.Embedding(max_aa_types * max_aa_types, feat_dim)
        self.relpos_embed = nn.Embedding(2 * max_relpos + 1, feat_dim)
        self.aapair_to_distcoef = nn.Embedding(max_aa_types * max_aa_types, max_num_atoms * max_num_atoms)
        nn.init.zeros_(self.aapair_to_distcoef.weight)
        self.distance_embed = nn.Sequential(nn.Linear(max_num_atoms * max_num_atoms, feat_dim), nn.ReLU(),
                                            nn.Linear(feat_dim, feat_dim), nn.ReLU())
        self.dihedral_embed = AngularEncoding()
        infeat_dim = 3 * feat_dim + self.dihedral_embed.get_out_dim(2)
        self.out_mlp = nn.Sequential(nn.Linear(infeat_dim, feat_dim), nn.ReLU(), nn.Linear(feat_dim, feat_dim), nn.ReLU(),
                                     nn.Linear(feat_dim, feat_dim))

    def native(self):
        return _native_embed(self, 'pair')

    def forward(self, aa, res_nb, chain_nb, pos_atoms, mask_atoms, structure_mask=None, sequence_mask=None):
        """aa, res_nb, chain_nb (N,L); pos_atoms (N,L,A,3); mask_atoms (N,L,A); structure_mask, sequence_mask (N,L) or None
        -> (N,L,L,feat_dim).  pair.py:37-101."""
        _native.forbid_training_graph(self, 'PairEmbedding.forward')
        with torch.no_grad():
            return self._forward(aa, res_nb, chain_nb, pos_atoms, mask_atoms, structure_mask, sequence_mask)

    def _forward(self, aa, res_nb, chain_nb, pos_atoms, mask_atoms, structure_mask, sequence_mask):
        nm = self.native()
        aa, res_nb, chain_nb, pos, mask, sm, qm, N, L, A = _common_inputs(aa, res_nb, chain_nb, pos_atoms, mask_atoms,
                                                                          structure_mask, sequence_mask)
        out = torch.empty(N, L, L, 64, device=aa.device, dtype=torch.float32)
        _capi.check(_capi.lib().abopt_pair_embed_forward(nm.handle, N, L, A, _capi.ptr(aa), _capi.ptr(res_nb), _capi.ptr(chain_nb),
                                                         _capi.ptr(pos), _capi.ptr(mask), _capi.ptr(sm), _capi.ptr(qm), _capi.ptr(out),
                                                         _capi.stream_ptr(aa.device)))
        return out


class ResidueEmbedding(nn.Module):
    """ResidueEmbedding with the reference's constructor, state-dict keys and call signature
    (/root/reference/AbDock/src/modules/encoders/residue.py:9-94), executing as one sm_100a kernel (csrc/k_res_embed.cu)."""

    def __init__(self, feat_dim, max_num_atoms, max_aa_types=22):
        super().__init__()
        if (feat_dim, max_aa_types) != (128, 22) or not 4 <= max_num_atoms <= 15:
            raise ValueError('the sm_100a kernel is specialised for the reference configuration: feat_dim 128, 22 amino-acid '
                             'types, 4..15 atoms per residue')
        self.max_num_atoms, self.max_aa_types = max_num_atoms, max_aa_types
        self.aatype_embed = nn.Embedding(max_aa_types, feat_dim)
        self.dihed_embed = AngularEncoding()
        self.type_embed = nn.Embedding(10, feat_dim, padding_idx=0)
        infeat_dim = feat_dim + max_aa_types * max_num_atoms * 3 + self.dihed_embed.get_out_dim(3) + feat_dim
        self.mlp = nn.Sequential(nn.Linear(infeat_dim, feat_dim * 2), nn.ReLU(), nn.Linear(feat_dim * 2, feat_dim), nn.ReLU(),
                                 nn.Linear(feat_dim, feat_dim), nn.ReLU(), nn.Linear(feat_dim, feat_dim))

    def native(self):
        return _native_embed(self, 'res')

    def forward(self, aa, res_nb, chain_nb, pos_atoms, mask_atoms, fragment_type, structure_mask=None, sequence_mask=None):
        """-> (N, L, feat_dim).  residue.py:27-94."""
        _native.forbid_training_graph(self, 'ResidueEmbedding.forward')
        with torch.no_grad():
            return self._forward(aa, res_nb, chain_nb, pos_atoms, mask_atoms, fragment_type, structure_mask, sequence_mask)

    def _forward(self, aa, res_nb, chain_nb, pos_atoms, mask_atoms, fragment_type, structure_mask, sequence_mask):
        nm = self.native()
        aa, res_nb, chain_nb, pos, mask, sm, qm, N, L, A = _common_inputs(aa, res_nb, chain_nb, pos_atoms, mask_atoms,
                                                                          structure_mask, sequence_mask)
        ft = _capi.cuda_i64(fragment_type, 'fragment_type')
        if ft.shape != (N, L):
            raise ValueError(f'bad shape: fragment_type {tuple(ft.shape)}')
        out = torch.empty(N, L, 128, device=aa.device, dtype=torch.float32)
        _capi.check(_capi.lib().abopt_res_embed_forward(nm.handle, N, L, A, _capi.ptr(aa), _capi.ptr(res_nb), _capi.ptr(chain_nb),
                                                        _capi.ptr(pos), _capi.ptr(mask), _capi.ptr(ft), _capi.ptr(sm), _capi.ptr(qm),
                                                        _capi.ptr(out), _capi.stream_ptr(aa.device)))
        return out
