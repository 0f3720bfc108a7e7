"""Oracle: invariant-point-attention block / encoder (test infrastructure only).

Restates /root/reference/AbDock/src/modules/encoders/ga.py (byte-identical to the AbDesign
copy modulo imports) as pure functions over a flat state-dict `W` whose keys carry `prefix`
(e.g. 'eps_net.encoder.blocks.0.').  Shapes: N complexes, L residues, F=128 node channels,
C=64 pair channels, H=12 heads, D=32 qk/value channels, P=8 points.

`materialize=True` reproduces the reference's broadcast-multiply-then-sum evaluation order
(and its multi-GB temporaries) so that timing this port on the CPU is representative of
the reference's own CPU path; `materialize=False` contracts with einsum (same maths, used
by tests at larger shapes).
"""
import math

import torch
import torch.nn.functional as F

from .geometry import frame_to_global, frame_to_local, unit_vector

H, D, P = 12, 32, 8   # ga.py:42-43 defaults (num_heads, query_key_dim = value_dim, num points)


def layer_norm(x, gamma, beta, eps=1e-10):
    """Hand-rolled LN: biased variance, eps inside the sqrt.  common/layers.py:146-155."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / (var + eps).sqrt() * gamma + beta


def attention_weights(logits, mask, inf=1e5):
    """ga.py:11-26.  logits (N,L,L,H), mask (N,L) bool -> alpha (N,L,L,H)."""
    row = mask[:, :, None, None]
    pair = row & mask[:, None, :, None]
    logits = torch.where(pair, logits, logits - inf)
    alpha = torch.softmax(logits, dim=2)
    return torch.where(row, alpha, torch.zeros_like(alpha))


def block_logits(W, prefix, R, t, x, z, materialize=True):
    """Sum of node, pair and spatial logits BEFORE the sqrt(1/3) scale.  ga.py:81-112,159-164."""
    N, L, _ = x.shape
    q = F.linear(x, W[prefix + 'proj_query.weight']).view(N, L, H, D)
    k = F.linear(x, W[prefix + 'proj_key.weight']).view(N, L, H, D)
    if materialize:
        node = (q[:, :, None] * k[:, None] * (1 / math.sqrt(D))).sum(-1)          # ga.py:84-85
    else:
        node = torch.einsum('nihd,njhd->nijh', q, k) * (1 / math.sqrt(D))
    pair = F.linear(z, W[prefix + 'proj_pair_bias.weight'])                       # ga.py:89

    qp = F.linear(x, W[prefix + 'proj_query_point.weight']).view(N, L, H * P, 3)  # ga.py:96-99
    kp = F.linear(x, W[prefix + 'proj_key_point.weight']).view(N, L, H * P, 3)    # ga.py:102-105
    qg = frame_to_global(R, t, qp).reshape(N, L, H, P * 3)
    kg = frame_to_global(R, t, kp).reshape(N, L, H, P * 3)
    if materialize:
        d2 = ((qg[:, :, None] - kg[:, None]) ** 2).sum(-1)                        # ga.py:108
    else:
        # same direct (q-k)^2 arithmetic, evaluated in row chunks to bound the temporary
        d2 = torch.cat([((qg[:, i0:i0 + 16, None] - kg[:, None]) ** 2).sum(-1)
                        for i0 in range(0, L, 16)], dim=1)
    gamma = F.softplus(W[prefix + 'spatial_coef'])                                # (1,1,1,H)
    spatial = d2 * ((-1 * gamma * math.sqrt(2 / (9 * P))) / 2)                    # ga.py:109-111
    return node + pair + spatial


def block_aggregate(W, prefix, alpha, R, t, x, z, materialize=True):
    """Pair, node and point aggregation -> (N, L, 1824).  ga.py:114-147,168-174."""
    N, L, _ = x.shape
    if materialize:
        p2n = (alpha[..., None] * z[:, :, :, None, :]).sum(2)                     # ga.py:116-117
    else:
        p2n = torch.einsum('nijh,nijc->nihc', alpha, z)
    v = F.linear(x, W[prefix + 'proj_value.weight']).view(N, L, H, D)             # ga.py:122
    if materialize:
        node = (alpha[..., None] * v[:, None]).sum(2)                             # ga.py:123-124
    else:
        node = torch.einsum('nijh,njhd->nihd', alpha, v)
    vp = F.linear(x, W[prefix + 'proj_value_point.weight']).view(N, L, H, P, 3)
    vg = frame_to_global(R, t, vp)                                                # ga.py:131-132
    if materialize:
        agg = (alpha.reshape(N, L, L, H, 1, 1) * vg[:, None]).sum(2)              # ga.py:133-135
    else:
        agg = torch.einsum('nijh,njhpc->nihpc', alpha, vg)
    pts = frame_to_local(R, t, agg)                                               # ga.py:137
    dist = pts.norm(dim=-1)                                                       # ga.py:138
    dirn = unit_vector(pts, dim=-1, eps=1e-4)                                     # ga.py:139
    return torch.cat([p2n.reshape(N, L, -1), node.reshape(N, L, -1), pts.reshape(N, L, -1),
                      dist.reshape(N, L, -1), dirn.reshape(N, L, -1)], -1)


def block_tail(W, prefix, x, feat, mask):
    """out_transform -> mask -> LN(x + .) -> 3-layer ReLU MLP -> LN.  ga.py:173-178."""
    y = F.linear(feat, W[prefix + 'out_transform.weight'], W[prefix + 'out_transform.bias'])
    y = torch.where(mask[..., None], y, torch.zeros_like(y))                       # layers.py:6-7
    h = layer_norm(x + y, W[prefix + 'layer_norm_1.gamma'], W[prefix + 'layer_norm_1.beta'])
    m = F.relu(F.linear(h, W[prefix + 'mlp_transition.0.weight'], W[prefix + 'mlp_transition.0.bias']))
    m = F.relu(F.linear(m, W[prefix + 'mlp_transition.2.weight'], W[prefix + 'mlp_transition.2.bias']))
    m = F.linear(m, W[prefix + 'mlp_transition.4.weight'], W[prefix + 'mlp_transition.4.bias'])
    return layer_norm(h + m, W[prefix + 'layer_norm_2.gamma'], W[prefix + 'layer_norm_2.beta'])


def ga_block(W, prefix, R, t, x, z, mask, materialize=True, return_parts=False):
    """GABlock.forward.  ga.py:149-178."""
    logits = block_logits(W, prefix, R, t, x, z, materialize)
    alpha = attention_weights(logits * math.sqrt(1 / 3), mask)                     # ga.py:166
    feat = block_aggregate(W, prefix, alpha, R, t, x, z, materialize)
    out = block_tail(W, prefix, x, feat, mask)
    if return_parts:
        return out, dict(logits=logits, alpha=alpha, feat=feat)
    return out


def ga_encoder(W, prefix, R, t, x, z, mask, num_layers, materialize=True):
    """GAEncoder.forward: `num_layers` sequential blocks over the same R, t, z, mask.  ga.py:190-193."""
    for l in range(num_layers):
        x = ga_block(W, f'{prefix}blocks.{l}.', R, t, x, z, mask, materialize)
    return x
