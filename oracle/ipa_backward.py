"""Oracle (TEST INFRASTRUCTURE ONLY): the backward pass of one GABlock written out by hand -- the formulas the CUDA backward
kernels will implement, restated without autograd so that every intermediate the kernels need
(what must be kept from the forward pass, what can be recomputed, which contractions appear) is explicit.

The reference has no hand-written backward: `loss.backward()` in train.py differentiates
/root/reference/AbDock/src/modules/encoders/ga.py:149-178 through torch autograd.  tests/test_oracle_golden.py checks this file
against autograd of oracle.ipa.ga_block (itself pinned to the reference), in fp64 to 1e-9 and in fp32.

Frames (R, t) and the mask carry no gradient: in a training step they come from the noised inputs (dpm_full.py:162-167).
Notation as oracle/ipa.py: N complexes, L residues, F=128, C=64, H=12, D=32, P=8.
"""
import math

import torch
import torch.nn.functional as F

from .ipa import H, D, P, attention_weights, block_aggregate, block_logits, layer_norm


def _layer_norm_backward(g, s, gamma, eps=1e-10):
    """y = (s - mu) / sqrt(var + eps) * gamma + beta (layers.py:146-155) -> d s, d gamma, d beta."""
    mu = s.mean(-1, keepdim=True)
    inv = 1.0 / ((s - mu).pow(2).mean(-1, keepdim=True) + eps).sqrt()
    xh = (s - mu) * inv
    gx = g * gamma
    ds = inv * (gx - gx.mean(-1, keepdim=True) - xh * (gx * xh).mean(-1, keepdim=True))
    red = tuple(range(g.dim() - 1))
    return ds, (g * xh).sum(red), g.sum(red)


def _linear_backward(g, inp, weight):
    """y = inp W^T (+ b) -> d inp, d W, d b."""
    g2, i2 = g.reshape(-1, g.shape[-1]), inp.reshape(-1, inp.shape[-1])
    return g @ weight, g2.t() @ i2, g2.sum(0)


def ga_block_backward(W, prefix, R, t, x, z, mask, g_out):
    """d loss / d (x, z, every weight of the block) given g_out = d loss / d GABlock(x) (N,L,F).

    Returns (g_x, g_z, {state-dict key: gradient}).  Kept from the forward pass: x (the block input).  Recomputed here:
    projections, logits, alpha, the aggregate, the tail's activations -- the recompute-based scheme of DESIGN.md section 7."""
    N, L, Fd = x.shape
    C = z.shape[-1]
    w = lambda name: W[prefix + name]
    grads = {}
    # ------------------------------------------------------------------ forward recompute (ga.py:149-178)
    q = F.linear(x, w('proj_query.weight')).view(N, L, H, D)
    k = F.linear(x, w('proj_key.weight')).view(N, L, H, D)
    v = F.linear(x, w('proj_value.weight')).view(N, L, H, D)
    to_g = lambda p: torch.einsum('nlab,nlhpb->nlhpa', R, p) + t[:, :, None, None, :]            # R p + t
    qg = to_g(F.linear(x, w('proj_query_point.weight')).view(N, L, H, P, 3))
    kg = to_g(F.linear(x, w('proj_key_point.weight')).view(N, L, H, P, 3))
    vg = to_g(F.linear(x, w('proj_value_point.weight')).view(N, L, H, P, 3))
    cP = math.sqrt(2 / (9 * P)) / 2
    coef = -F.softplus(w('spatial_coef')).reshape(H) * cP                                         # ga.py:109-111
    diff = qg[:, :, None] - kg[:, None]                                                           # (N,i,j,H,P,3)
    d2 = diff.pow(2).sum((-1, -2))
    scale = math.sqrt(1 / 3)
    logits = (torch.einsum('nihd,njhd->nijh', q, k) / math.sqrt(D) + F.linear(z, w('proj_pair_bias.weight')) + d2 * coef) * scale
    alpha = attention_weights(logits, mask)                                                       # ga.py:11-26
    agg = torch.einsum('nijh,njhpc->nihpc', alpha, vg)
    pts = torch.einsum('nlba,nlhpb->nlhpa', R, agg - t[:, :, None, None, :])                      # R^T (agg - t)
    nrm = pts.norm(dim=-1, keepdim=True)
    feat = torch.cat([torch.einsum('nijh,nijc->nihc', alpha, z).reshape(N, L, -1), torch.einsum('nijh,njhd->nihd', alpha, v).reshape(N, L, -1),
                      pts.reshape(N, L, -1), nrm.reshape(N, L, -1), (pts / (nrm + 1e-4)).reshape(N, L, -1)], -1)
    y = F.linear(feat, w('out_transform.weight'), w('out_transform.bias')) * mask[..., None]
    s1 = x + y
    h = layer_norm(s1, w('layer_norm_1.gamma'), w('layer_norm_1.beta'))
    a0 = F.linear(h, w('mlp_transition.0.weight'), w('mlp_transition.0.bias'))
    a1 = F.linear(F.relu(a0), w('mlp_transition.2.weight'), w('mlp_transition.2.bias'))
    m3 = F.linear(F.relu(a1), w('mlp_transition.4.weight'), w('mlp_transition.4.bias'))
    s2 = h + m3
    # ------------------------------------------------------------------ tail backward (ga.py:173-178)
    g_s2, grads['layer_norm_2.gamma'], grads['layer_norm_2.beta'] = _layer_norm_backward(g_out, s2, w('layer_norm_2.gamma'))
    g_r1, grads['mlp_transition.4.weight'], grads['mlp_transition.4.bias'] = _linear_backward(g_s2, F.relu(a1), w('mlp_transition.4.weight'))
    g_r0, grads['mlp_transition.2.weight'], grads['mlp_transition.2.bias'] = _linear_backward(g_r1 * (a1 > 0), F.relu(a0), w('mlp_transition.2.weight'))
    g_h, grads['mlp_transition.0.weight'], grads['mlp_transition.0.bias'] = _linear_backward(g_r0 * (a0 > 0), h, w('mlp_transition.0.weight'))
    g_s1, grads['layer_norm_1.gamma'], grads['layer_norm_1.beta'] = _layer_norm_backward(g_h + g_s2, s1, w('layer_norm_1.gamma'))
    g_x = g_s1.clone()
    g_feat, grads['out_transform.weight'], grads['out_transform.bias'] = _linear_backward(g_s1 * mask[..., None], feat, w('out_transform.weight'))
    # ------------------------------------------------------------------ aggregate backward (ga.py:114-147)
    o0, o1, o2, o3 = H * C, H * C + H * D, H * C + H * D + H * P * 3, H * C + H * D + H * P * 4
    g_p2n = g_feat[..., :o0].reshape(N, L, H, C)
    g_node = g_feat[..., o0:o1].reshape(N, L, H, D)
    g_pts = g_feat[..., o1:o2].reshape(N, L, H, P, 3).clone()
    g_dist = g_feat[..., o2:o3].reshape(N, L, H, P, 1)
    g_dir = g_feat[..., o3:].reshape(N, L, H, P, 3)
    unit = pts / nrm.clamp_min(1e-30)
    g_pts += g_dist * unit                                                                        # d |p| = p / |p|
    g_pts += g_dir / (nrm + 1e-4) - unit * (g_dir * pts).sum(-1, keepdim=True) / (nrm + 1e-4).pow(2)   # d (p / (|p| + eps))
    g_agg = torch.einsum('nlab,nlhpb->nlhpa', R, g_pts)                                           # pts = R^T (agg - t)
    g_alpha = (torch.einsum('nihc,nijc->nijh', g_p2n, z) + torch.einsum('nihd,njhd->nijh', g_node, v)
               + torch.einsum('nihpc,njhpc->nijh', g_agg, vg))
    g_z = torch.einsum('nijh,nihc->nijc', alpha, g_p2n)
    g_v = torch.einsum('nijh,nihd->njhd', alpha, g_node)
    g_vg = torch.einsum('nijh,nihpc->njhpc', alpha, g_agg)
    # ------------------------------------------------------------------ softmax and logits backward (ga.py:11-26, 81-112, 159-166)
    g_log = alpha * (g_alpha - (alpha * g_alpha).sum(2, keepdim=True)) * scale                    # masked rows / keys have alpha = 0
    g_q = torch.einsum('nijh,njhd->nihd', g_log, k) / math.sqrt(D)
    g_k = torch.einsum('nijh,nihd->njhd', g_log, q) / math.sqrt(D)
    g_z += g_log @ w('proj_pair_bias.weight')                                                     # (N,i,j,H) x (H,C)
    grads['proj_pair_bias.weight'] = torch.einsum('nijh,nijc->hc', g_log, z)
    g_coef = (g_log * d2).sum((0, 1, 2))
    grads['spatial_coef'] = (g_coef * (-cP) * torch.sigmoid(w('spatial_coef').reshape(H))).reshape(1, 1, 1, H)
    g_d = 2 * (g_log * coef)[..., None, None] * diff                                              # d (d2 coef) / d (qg - kg)
    g_qg, g_kg = g_d.sum(2), -g_d.sum(1)
    # ------------------------------------------------------------------ projections backward
    to_l = lambda g: torch.einsum('nlba,nlhpb->nlhpa', R, g).reshape(N, L, -1)                    # q_global = R q_local + t
    for name, g in (('proj_query', g_q.reshape(N, L, -1)), ('proj_key', g_k.reshape(N, L, -1)), ('proj_value', g_v.reshape(N, L, -1)),
                    ('proj_query_point', to_l(g_qg)), ('proj_key_point', to_l(g_kg)), ('proj_value_point', to_l(g_vg))):
        gx, grads[name + '.weight'], _ = _linear_backward(g, x, w(name + '.weight'))
        g_x += gx
    return g_x, g_z, {prefix + k: v for k, v in grads.items()}
