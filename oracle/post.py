"""Oracle (TEST INFRASTRUCTURE ONLY -- imported by tests/ and the CPU legs of bench.py, never by the product path):
CPU restatement of the steps right after the sampling loop, SURVEY.md section 8f rank 3.

    reconstruct_backbone_partially   /root/reference/AbDock/src/modules/common/geometry.py:450-480
    reconstruct_backbone             geometry.py:404-447  (local_to_global :72-92, compose_chain :120-140,
                                     get_backbone_dihedral_angles :307-348, topology.py:5-24)
    calc_per_rmsd / calc_avg_rmsd / rank_commoness
                                     /root/reference/AbDock/src/tools/runner/design_for_testset.py:556-589

The ideal backbone geometry (`backbone_atom_coordinates_tensor` (21,3,3), `bb_oxygen_coordinate_tensor` (21,3),
utils/protein/constants.py:310-320) is DATA of the reference and is passed in by the caller; tests use the copy stored in the
golden fixture (written by tests/golden/make_golden.py from the reference module) or random tables.

Pinned against the unmodified reference by tests/test_oracle_vs_reference.py and tests/golden/post_loop.npz.
"""
import torch

from .pair_embed import _dihedral, ATOM_CA


def reconstruct_backbone(R, t, aa, chain_nb, res_nb, mask, bb_table, o_table):
    """geometry.py:404-447 -> (N, L, 4, 3): ideal N, CA, C placed by the residue frame, O placed after turning about the
    CA-C axis (the frame's x axis) by psi."""
    aa = aa.clamp(0, 20)
    ideal = bb_table.to(t)[aa]                                                            # (N,L,3,3)
    nca_c = torch.einsum('nlij,nlaj->nlai', R, ideal) + t[:, :, None, :]                    # local_to_global
    n, ca, c = nca_c[:, :, 0], nca_c[:, :, 1], nca_c[:, :, 2]
    bonded = ((res_nb[:, 1:] - res_nb[:, :-1]).abs() == 1) & (chain_nb[:, 1:] == chain_nb[:, :-1]) & mask[:, :-1]
    psi = torch.zeros_like(t[:, :, 0])
    psi[:, :-1] = _dihedral(n[:, :-1], ca[:, :-1], c[:, :-1], n[:, 1:]) * bonded            # geometry.py:424-425
    cs, sn = torch.cos(psi), torch.sin(psi)
    zero, one = torch.zeros_like(cs), torch.ones_like(cs)
    turn = torch.stack([one, zero, zero, zero, cs, -sn, zero, sn, cs], dim=-1).reshape(*psi.shape, 3, 3)
    R_o = torch.matmul(R, turn)                                                           # compose_chain: R1 R2, t unchanged
    o = torch.einsum('nlij,nlj->nli', R_o, o_table.to(t)[aa]) + t
    return torch.cat([nca_c, o[:, :, None, :]], dim=2)


def reconstruct_backbone_partially(pos_ctx, R_new, t_new, aa, chain_nb, res_nb, mask_atoms, mask_recons, bb_table, o_table):
    """geometry.py:450-480 -> pos_new (N,L,A,3), mask_new (N,L,A)."""
    N, L, A = mask_atoms.shape
    rebuilt = torch.zeros_like(pos_ctx)
    rebuilt[:, :, :4] = reconstruct_backbone(R_new, t_new, aa, chain_nb, res_nb, mask_atoms[:, :, ATOM_CA], bb_table, o_table)
    bb_only = torch.zeros_like(mask_atoms)
    bb_only[:, :, :4] = True
    sel = mask_recons[:, :, None]
    return torch.where(sel[..., None], rebuilt, pos_ctx), torch.where(sel, bb_only, mask_atoms)


def pairwise_rmsd(structures):
    """calc_per_rmsd, design_for_testset.py:556-563: (B, M, 3) -> (B, B)."""
    diff = structures[:, None] - structures[None, :]
    return torch.sqrt((diff ** 2).sum(-1).mean(-1))


def average_rmsd(structures):
    """calc_avg_rmsd, design_for_testset.py:566-570."""
    B = structures.shape[0]
    return pairwise_rmsd(structures).sum() / (B * (B - 1))


def commonness(structures):
    """Mean RMSD of every structure to the others (the score rank_commoness sorts, design_for_testset.py:585-586)."""
    return pairwise_rmsd(structures).sum(-1) / (structures.shape[0] - 1)


def rank_commonness(structures, k):
    """rank_commoness, design_for_testset.py:573-589: indices of the k structures with the smallest mean RMSD, best first."""
    return torch.topk(commonness(structures), k=k, largest=False)[1]
