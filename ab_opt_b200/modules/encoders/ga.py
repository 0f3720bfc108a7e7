"""GABlock / GAEncoder with the reference's constructor signatures and state-dict keys
(/root/reference/AbDock/src/modules/encoders/ga.py:39-193), executing on the sm_100a kernels of
libabopt_b200 through the C ABI.  No PyTorch arithmetic happens here."""
import numpy as np
import torch
import torch.nn as nn

from .. import _native
from ... import _capi
from ..common.layers import LayerNorm


class GABlock(_native.NativeOwner, nn.Module):
    _native_scope = _capi.SCOPE_ENCODER

    def __init__(self, node_feat_dim, pair_feat_dim, value_dim=32, query_key_dim=32, num_query_points=8,
                 num_value_points=8, num_heads=12, bias=False):
        super().__init__()
        if (node_feat_dim, pair_feat_dim, value_dim, query_key_dim, num_query_points, num_value_points, num_heads,
                bias) != (128, 64, 32, 32, 8, 8, 12, False):
            raise ValueError('the sm_100a kernels are specialised for the reference configuration: node 128, pair 64, '
                             '12 heads x 32 channels, 8 query/value points, no projection bias')
        self.node_feat_dim, self.pair_feat_dim = node_feat_dim, pair_feat_dim
        self.value_dim, self.query_key_dim = value_dim, query_key_dim
        self.num_query_points, self.num_value_points, self.num_heads = num_query_points, num_value_points, num_heads
        self.proj_query = nn.Linear(node_feat_dim, query_key_dim * num_heads, bias=bias)
        self.proj_key = nn.Linear(node_feat_dim, query_key_dim * num_heads, bias=bias)
        self.proj_value = nn.Linear(node_feat_dim, value_dim * num_heads, bias=bias)
        self.proj_pair_bias = nn.Linear(pair_feat_dim, num_heads, bias=bias)
        self.spatial_coef = nn.Parameter(torch.full([1, 1, 1, num_heads], fill_value=float(np.log(np.exp(1.) - 1.))))
        self.proj_query_point = nn.Linear(node_feat_dim, num_query_points * num_heads * 3, bias=bias)
        self.proj_key_point = nn.Linear(node_feat_dim, num_query_points * num_heads * 3, bias=bias)
        self.proj_value_point = nn.Linear(node_feat_dim, num_value_points * num_heads * 3, bias=bias)
        self.out_transform = nn.Linear(num_heads * pair_feat_dim + num_heads * value_dim
                                       + num_heads * num_value_points * 7, node_feat_dim)
        self.layer_norm_1 = LayerNorm(node_feat_dim)
        self.mlp_transition = nn.Sequential(nn.Linear(node_feat_dim, node_feat_dim), nn.ReLU(),
                                            nn.Linear(node_feat_dim, node_feat_dim), nn.ReLU(),
                                            nn.Linear(node_feat_dim, node_feat_dim))
        self.layer_norm_2 = LayerNorm(node_feat_dim)

    # stand-alone use: a one-layer encoder handle
    def _native_config(self):
        return _capi.Config(1, 100, 0, 0, 0.0, 0.0, 0, _capi.SCOPE_ENCODER)

    def _native_state(self):
        return {'eps_net.encoder.blocks.0.' + k: v for k, v in self.state_dict(keep_vars=True).items()}

    @torch.no_grad()
    def forward(self, R, t, x, z, mask):
        """R (N,L,3,3), t (N,L,3), x (N,L,F), z (N,L,L,C), mask (N,L) -> (N,L,F).  ga.py:149-178."""
        return _run_block(self.native(), 0, R, t, x, z, mask)


def _prep(R, t, x, z, mask):
    R = _capi.cuda_f32(R, 'R'); t = _capi.cuda_f32(t, 't'); x = _capi.cuda_f32(x, 'x'); z = _capi.cuda_f32(z, 'z')
    mask = _capi.cuda_mask(mask, 'mask')
    N, L = mask.shape
    if R.shape != (N, L, 3, 3) or t.shape != (N, L, 3) or x.shape != (N, L, 128) or z.shape != (N, L, L, 64):
        raise ValueError(f'bad shapes: R {tuple(R.shape)} t {tuple(t.shape)} x {tuple(x.shape)} z {tuple(z.shape)} '
                         f'mask {tuple(mask.shape)}')
    return R, t, x, z, mask, N, L


def _run_block(nm, layer, R, t, x, z, mask):
    R, t, x, z, mask, N, L = _prep(R, t, x, z, mask)
    out = torch.empty_like(x)
    _capi.check(_capi.lib().abopt_ga_block_forward(nm.handle, layer, N, L, _capi.ptr(R), _capi.ptr(t), _capi.ptr(x),
                                                   _capi.ptr(z), _capi.ptr(mask), _capi.ptr(out),
                                                   _capi.stream_ptr(x.device)))
    return out


class GAEncoder(_native.NativeOwner, nn.Module):
    _native_scope = _capi.SCOPE_ENCODER

    def __init__(self, node_feat_dim, pair_feat_dim, num_layers, ga_block_opt={}):
        super().__init__()
        self.blocks = nn.ModuleList([GABlock(node_feat_dim, pair_feat_dim, **ga_block_opt) for _ in range(num_layers)])

    def _native_config(self):
        return _capi.Config(len(self.blocks), 100, 0, 0, 0.0, 0.0, 0, _capi.SCOPE_ENCODER)

    def _native_state(self):
        return {'eps_net.encoder.' + k: v for k, v in self.state_dict(keep_vars=True).items()}

    @torch.no_grad()
    def forward(self, R, t, res_feat, pair_feat, mask):
        """All blocks over the same R, t, pair_feat, mask.  ga.py:190-193."""
        nm = self.native()
        R, t, x, z, mask, N, L = _prep(R, t, res_feat, pair_feat, mask)
        out = torch.empty_like(x)
        _capi.check(_capi.lib().abopt_ga_encoder_forward(nm.handle, N, L, _capi.ptr(R), _capi.ptr(t), _capi.ptr(x),
                                                         _capi.ptr(z), _capi.ptr(mask), _capi.ptr(out),
                                                         _capi.stream_ptr(x.device)))
        return out

    @torch.no_grad()
    def block_taps(self, layer, R, t, x, z, mask):
        """Parity taps of one block: (alpha (N,L,L,12), aggregate (N,L,1824)) -- ga.py:166-174."""
        nm = self.native()
        R, t, x, z, mask, N, L = _prep(R, t, x, z, mask)
        alpha = torch.empty(N, L, L, 12, device=x.device)
        feat = torch.empty(N, L, 1824, device=x.device)
        _capi.check(_capi.lib().abopt_ga_block_taps(nm.handle, layer, N, L, _capi.ptr(R), _capi.ptr(t), _capi.ptr(x),
                                                    _capi.ptr(z), _capi.ptr(mask), _capi.ptr(alpha), _capi.ptr(feat),
                                                    _capi.stream_ptr(x.device)))
        return alpha, feat

    @torch.no_grad()
    def block_backward(self, layer, R, t, x, z, mask, g_out):
        """Backward of one GABlock (ga.py:149-178 under autograd; csrc/k_backward.cu): g_out = d loss / d block output ->
        (d loss / d x, d loss / d z, {state-dict key of the block: gradient})."""
        nm = self.native()
        R, t, x, z, mask, N, L = _prep(R, t, x, z, mask)
        g_out = _capi.cuda_f32(g_out, 'g_out')
        g_x, g_z = torch.empty_like(x), torch.empty_like(z)
        st = _capi.stream_ptr(x.device)
        _capi.check(_capi.lib().abopt_ga_block_backward(nm.handle, layer, N, L, _capi.ptr(R), _capi.ptr(t), _capi.ptr(x), _capi.ptr(z),
                                                        _capi.ptr(mask), _capi.ptr(g_out), _capi.ptr(g_x), _capi.ptr(g_z), st))
        grads = {}
        pre = f'blocks.{layer}.'
        for k, v in self.state_dict().items():
            if k.startswith(pre):
                g = torch.empty(v.numel(), device=x.device)
                _capi.check(_capi.lib().abopt_model_get_grad(nm.handle, ('eps_net.encoder.' + k).encode(), _capi.ptr(g), v.numel(), st))
                grads[k] = g.view(v.shape)
        return g_x, g_z, grads
