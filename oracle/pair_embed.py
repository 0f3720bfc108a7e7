"""Oracle (TEST INFRASTRUCTURE ONLY -- imported by tests/, smoke() and bench.py's CPU legs, never by the product path):
CPU restatement of the reference's O(L^2) pair featurisation, SURVEY.md section 8f rank 1.

    PairEmbedding.forward          /root/reference/AbDock/src/modules/encoders/pair.py:37-101  (AbDesign: same file, same lines)
    pairwise_dihedrals             /root/reference/AbDock/src/modules/common/geometry.py:351-376
    dihedral_from_four_points      geometry.py:254-271
    AngularEncoding.forward        /root/reference/AbDock/src/modules/common/layers.py:85-106

Pinned against the unmodified reference by tests/test_oracle_vs_reference.py (build container) and by the fixture
tests/golden/pair_embed.npz (tests/golden/make_golden.py ran the reference class).

Works in whatever floating dtype the weights and coordinates come in (fp32 for parity, fp64 as the arbiter).
"""
import math

import numpy as np
import torch

MAX_AA, MAX_RELPOS, UNK = 22, 32, 20        # pair.py:12 defaults; constants.py:108 (AA.UNK)
ATOM_N, ATOM_CA, ATOM_C = 0, 1, 2           # constants.py:139-140 (BBHeavyAtom)
C_DIM = 64


def make_state_dict(seed=0, num_atoms=15, feat_dim=C_DIM):
    """Seeded synthetic PairEmbedding.state_dict() with the reference's keys and shapes (pair.py:12-35).
    `aapair_to_distcoef` is zero-initialised in the reference (pair.py:21); trained values are not, so it is drawn."""
    rs = np.random.RandomState(seed)
    A2 = num_atoms * num_atoms
    f32 = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32))
    W = {
        'aa_pair_embed.weight': f32(rs.standard_normal((MAX_AA * MAX_AA, feat_dim))),
        'relpos_embed.weight': f32(rs.standard_normal((2 * MAX_RELPOS + 1, feat_dim))),
        'aapair_to_distcoef.weight': f32(0.7 * rs.standard_normal((MAX_AA * MAX_AA, A2))),
        'dihedral_embed.freq_bands': f32([1, 2, 3, 1.0, 1.0 / 2, 1.0 / 3]),
    }

    def lin(name, n_out, n_in):
        b = 1.0 / math.sqrt(n_in)
        W[name + '.weight'] = f32(rs.uniform(-b, b, (n_out, n_in)))
        W[name + '.bias'] = f32(rs.uniform(-b, b, (n_out,)))
    lin('distance_embed.0', feat_dim, A2)
    lin('distance_embed.2', feat_dim, feat_dim)
    lin('out_mlp.0', feat_dim, 3 * feat_dim + 26)
    lin('out_mlp.2', feat_dim, feat_dim)
    lin('out_mlp.4', feat_dim, feat_dim)
    return W


def synthetic_complex(seed, N, L, num_atoms_in=15, ragged=True, dtype=torch.float32):
    """Seeded protein-like inputs: a noisy helix-ish backbone per chain plus side-chain atoms, two or three chains,
    per-residue atom masks (missing side-chain atoms), trailing padding when `ragged`."""
    g = torch.Generator().manual_seed(seed)
    aa = torch.randint(0, 21, (N, L), generator=g)
    chain_nb = torch.zeros(N, L, dtype=torch.long)
    res_nb = torch.zeros(N, L, dtype=torch.long)
    for n in range(N):
        cuts = sorted(torch.randint(1, max(L, 2), (2,), generator=g).tolist())
        chain_nb[n, cuts[0]:] += 1
        chain_nb[n, cuts[1]:] += 1
        start = 0
        for c in range(3):
            idx = (chain_nb[n] == c).nonzero().flatten()
            if len(idx):
                gaps = (torch.rand(len(idx), generator=g) < 0.05).long() * torch.randint(1, 60, (len(idx),), generator=g)
                res_nb[n, idx] = 1 + torch.arange(len(idx)) + torch.cumsum(gaps, 0)
    ca = torch.cumsum(torch.randn(N, L, 3, generator=g) * 2.2, dim=1)                       # ~3.8 A steps
    pos = ca[:, :, None, :] + 1.6 * torch.randn(N, L, num_atoms_in, 3, generator=g)
    pos[:, :, ATOM_CA] = ca
    mask_atoms = torch.rand(N, L, num_atoms_in, generator=g) < 0.7
    mask_atoms[:, :, :4] = True
    if ragged:
        for n in range(N):
            ln = int(torch.randint(max(1, (3 * L) // 4), L + 1, (1,), generator=g))
            mask_atoms[n, ln:] = False
            pos[n, ln:] = 0.0                                                                # padding rows are all-zero
        mask_atoms[0, L // 3, ATOM_CA] = False                                               # a residue without CA
    context = torch.ones(N, L, dtype=torch.bool)
    context[:, L // 2: L // 2 + max(1, L // 8)] = False                                      # the generated stretch
    context &= mask_atoms[:, :, ATOM_CA]
    return dict(aa=aa, res_nb=res_nb, chain_nb=chain_nb, pos_atoms=pos.to(dtype), mask_atoms=mask_atoms, context_mask=context)


def _dihedral(p0, p1, p2, p3):
    """geometry.py:254-271: signed angle between the planes (p0,p1,p2) and (p1,p2,p3); cosine clamped to +-0.999999;
    NaN (degenerate normals) -> 0."""
    b_mid, b_prev, b_next = p2 - p1, p0 - p1, p3 - p2
    n_a = torch.linalg.cross(b_mid, b_prev, dim=-1)
    n_a = n_a / torch.linalg.norm(n_a, dim=-1, keepdim=True)
    n_b = torch.linalg.cross(b_mid, b_next, dim=-1)
    n_b = n_b / torch.linalg.norm(n_b, dim=-1, keepdim=True)
    handed = torch.sign((torch.linalg.cross(b_prev, b_next, dim=-1) * b_mid).sum(-1))
    ang = handed * torch.acos((n_a * n_b).sum(-1).clamp(-0.999999, 0.999999))
    return torch.nan_to_num(ang)


def inter_residue_dihedrals(pos_atoms):
    """geometry.py:351-376: phi_ij = dihedral(C_i, N_j, CA_j, C_j), psi_ij = dihedral(N_i, CA_i, C_i, N_j) -> (N,L,L,2)."""
    N, L = pos_atoms.shape[:2]
    n, ca, c = pos_atoms[:, :, ATOM_N], pos_atoms[:, :, ATOM_CA], pos_atoms[:, :, ATOM_C]
    row = lambda x: x[:, :, None, :].expand(N, L, L, 3)      # indexed by i
    col = lambda x: x[:, None, :, :].expand(N, L, L, 3)      # indexed by j
    phi = _dihedral(row(c), col(n), col(ca), col(c))
    psi = _dihedral(row(n), row(ca), row(c), col(n))
    return torch.stack([phi, psi], dim=-1)


def angular_encoding(x, freq_bands):
    """layers.py:97-106: per angle [x, sin(x f_1..6), cos(x f_1..6)] -> 13 numbers, angles concatenated."""
    xe = x.unsqueeze(-1)
    code = torch.cat([xe, torch.sin(xe * freq_bands), torch.cos(xe * freq_bands)], dim=-1)
    return code.reshape(*x.shape[:-1], -1)


def pair_embedding(W, aa, res_nb, chain_nb, pos_atoms, mask_atoms, structure_mask=None, sequence_mask=None):
    """PairEmbedding.forward, pair.py:37-101 -> (N, L, L, 64)."""
    lin = lambda name, x: torch.nn.functional.linear(x, W[name + '.weight'], W[name + '.bias'])
    A2 = W['aapair_to_distcoef.weight'].shape[1]
    A = int(round(math.sqrt(A2)))
    N, L = aa.shape
    pos_atoms, mask_atoms = pos_atoms[:, :, :A], mask_atoms[:, :, :A]                              # pair.py:54-55
    has_ca = mask_atoms[:, :, ATOM_CA]
    pair_ok = has_ca[:, :, None] & has_ca[:, None, :]                                              # pair.py:57-58
    if sequence_mask is not None:                                                                  # pair.py:62-64
        aa = torch.where(sequence_mask, aa, torch.full_like(aa, UNK))
    pair_type = aa[:, :, None] * MAX_AA + aa[:, None, :]                                           # pair.py:65
    f_type = W['aa_pair_embed.weight'][pair_type]
    offset = (res_nb[:, :, None] - res_nb[:, None, :]).clamp(-MAX_RELPOS, MAX_RELPOS) + MAX_RELPOS  # pair.py:70-73
    same_chain = chain_nb[:, :, None] == chain_nb[:, None, :]
    f_rel = W['relpos_embed.weight'][offset] * same_chain[..., None]                               # pair.py:74
    gap = pos_atoms[:, :, None, :, None, :] - pos_atoms[:, None, :, None, :, :]                    # (N,L,L,A,A,3)
    d_nm = (torch.linalg.norm(gap, dim=-1) / 10).reshape(N, L, L, A2)                              # pair.py:77-80
    coef = torch.nn.functional.softplus(W['aapair_to_distcoef.weight'][pair_type])                 # pair.py:81
    both = (mask_atoms[:, :, None, :, None] & mask_atoms[:, None, :, None, :]).reshape(N, L, L, A2)
    g = torch.exp(-coef * d_nm ** 2) * both                                                        # pair.py:82-84
    f_dist = torch.relu(lin('distance_embed.2', torch.relu(lin('distance_embed.0', g))))
    f_ang = angular_encoding(inter_residue_dihedrals(pos_atoms), W['dihedral_embed.freq_bands'])   # pair.py:90-91
    if structure_mask is not None:                                                                 # pair.py:85-87, 92-94
        keep = (structure_mask[:, :, None] & structure_mask[:, None, :])[..., None]
        f_dist, f_ang = f_dist * keep, f_ang * keep
    h = torch.cat([f_type, f_rel, f_dist, f_ang], dim=-1)                                          # pair.py:97
    h = torch.relu(lin('out_mlp.0', h))
    h = torch.relu(lin('out_mlp.2', h))
    h = lin('out_mlp.4', h)
    return h * pair_ok[..., None]                                                                  # pair.py:99


# ------------------------------------------------------------------------------------------ per-residue features
def make_residue_state_dict(seed=0, num_atoms=15, feat_dim=128):
    """Seeded synthetic ResidueEmbedding.state_dict() (encoders/residue.py:11-25): keys and shapes of the reference."""
    rs = np.random.RandomState(seed)
    f32 = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32))
    W = {'aatype_embed.weight': f32(rs.standard_normal((MAX_AA, feat_dim))),
         'dihed_embed.freq_bands': f32([1, 2, 3, 1.0, 1.0 / 2, 1.0 / 3]),
         'type_embed.weight': f32(rs.standard_normal((10, feat_dim)))}
    W['type_embed.weight'][0] = 0                                       # padding_idx=0 (residue.py:17)
    dims = [feat_dim + MAX_AA * num_atoms * 3 + 39 + feat_dim, 2 * feat_dim, feat_dim, feat_dim, feat_dim]
    for i in range(4):
        b = 1.0 / math.sqrt(dims[i])
        W[f'mlp.{2 * i}.weight'] = f32(rs.uniform(-b, b, (dims[i + 1], dims[i])))
        W[f'mlp.{2 * i}.bias'] = f32(rs.uniform(-b, b, (dims[i + 1],)))
    return W


def backbone_frames(ca, c, n):
    """construct_3d_basis, geometry.py:47-69: Gram-Schmidt on (C - CA, N - CA), columns e1 e2 e3; normalisation with +1e-6."""
    e1 = c - ca
    e1 = e1 / (torch.linalg.norm(e1, dim=-1, keepdim=True) + 1e-6)
    v2 = n - ca
    u2 = v2 - (e1 * v2).sum(-1, keepdim=True) * e1
    e2 = u2 / (torch.linalg.norm(u2, dim=-1, keepdim=True) + 1e-6)
    return torch.stack([e1, e2, torch.linalg.cross(e1, e2, dim=-1)], dim=-1)


def backbone_dihedrals(pos_atoms, chain_nb, res_nb, mask):
    """get_backbone_dihedral_angles, geometry.py:307-348 with topology.py:5-24: omega / phi need a bonded predecessor,
    psi a bonded successor; bonded(i, i+1) = |res_nb step| == 1, same chain, mask[i]."""
    n, ca, c = pos_atoms[:, :, ATOM_N], pos_atoms[:, :, ATOM_CA], pos_atoms[:, :, ATOM_C]
    bonded = ((res_nb[:, 1:] - res_nb[:, :-1]).abs() == 1) & (chain_nb[:, 1:] == chain_nb[:, :-1]) & mask[:, :-1]
    pad = torch.zeros_like(mask[:, :1])
    has_prev, has_next = torch.cat([pad, bonded], 1), torch.cat([bonded, pad], 1)
    zero = torch.zeros_like(ca[:, :1, 0])
    omega = torch.cat([zero, _dihedral(ca[:, :-1], c[:, :-1], n[:, 1:], ca[:, 1:])], 1)
    phi = torch.cat([zero, _dihedral(c[:, :-1], n[:, 1:], ca[:, 1:], c[:, 1:])], 1)
    psi = torch.cat([_dihedral(n[:, :-1], ca[:, :-1], c[:, :-1], n[:, 1:]), zero], 1)
    valid = torch.stack([has_prev, has_prev, has_next], -1)
    return torch.stack([omega, phi, psi], -1) * valid, valid


def residue_embedding(W, aa, res_nb, chain_nb, pos_atoms, mask_atoms, fragment_type, structure_mask=None, sequence_mask=None):
    """ResidueEmbedding.forward, encoders/residue.py:27-94 -> (N, L, 128)."""
    lin = lambda name, x: torch.nn.functional.linear(x, W[name + '.weight'], W[name + '.bias'])
    A = (W['mlp.0.weight'].shape[1] - 2 * W['aatype_embed.weight'].shape[1] - 39) // (3 * MAX_AA)
    N, L = aa.shape
    has_ca = mask_atoms[:, :, ATOM_CA]
    pos_atoms, mask_atoms = pos_atoms[:, :, :A], mask_atoms[:, :, :A]                              # residue.py:43-44
    if sequence_mask is not None:                                                                  # residue.py:47-49
        aa = torch.where(sequence_mask, aa, torch.full_like(aa, UNK))
    f_aa = W['aatype_embed.weight'][aa]
    ca = pos_atoms[:, :, ATOM_CA]
    R = backbone_frames(ca, pos_atoms[:, :, ATOM_C], pos_atoms[:, :, ATOM_N])                       # residue.py:53-58
    local = torch.einsum('nlkc,nlak->nlac', R, pos_atoms - ca[:, :, None, :])                       # R^T (x - t), geometry.py:94-113
    local = local * mask_atoms[..., None]                                                           # residue.py:60-61
    slot = torch.nn.functional.one_hot(aa, MAX_AA).to(local.dtype)                                  # residue.py:63-68
    f_crd = (slot[:, :, :, None, None] * local[:, :, None, :, :]).reshape(N, L, MAX_AA * A * 3)
    if structure_mask is not None:
        f_crd = f_crd * structure_mask[:, :, None]                                                  # residue.py:69-71
    ang, valid = backbone_dihedrals(pos_atoms, chain_nb, res_nb, has_ca)                            # residue.py:74
    f_ang = (angular_encoding(ang[..., None], W['dihed_embed.freq_bands']).reshape(N, L, 3, 13) * valid[..., None]).reshape(N, L, 39)
    if structure_mask is not None:                                                                  # residue.py:77-86
        near = structure_mask & torch.roll(structure_mask, 1, 1) & torch.roll(structure_mask, -1, 1)
        f_ang = f_ang * near[:, :, None]
    f_type = torch.nn.functional.embedding(fragment_type, W['type_embed.weight'], padding_idx=0)   # residue.py:17,89: row 0 gets no gradient
    h = torch.cat([f_aa, f_crd, f_ang, f_type], dim=-1)                                             # residue.py:89-92
    for i in range(3):
        h = torch.relu(lin(f'mlp.{2 * i}', h))
    return lin('mlp.6', h) * has_ca[:, :, None]                                                     # residue.py:93


def embedding_grads(which, W, g_out, *inputs):
    """Backward target of the featurisation for the training step: gradients of sum(out * g_out) with
    respect to every floating weight of `pair_embedding` (which='pair') or `residue_embedding` (which='residue'), by torch
    autograd through the oracle; g_out is d loss / d pair_feat (resp. res_feat) as the hand-written step in
    oracle/epsnet_backward.py returns it.  Pinned to the reference's own autograd by tests/golden/pair_embed.npz."""
    fn = pair_embedding if which == 'pair' else residue_embedding
    Wg = {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() and 'freq_bands' not in k else v) for k, v in W.items()}
    with torch.enable_grad():
        (fn(Wg, *inputs) * g_out).sum().backward()
    return {k: v.grad for k, v in Wg.items() if v.requires_grad and v.grad is not None}
