"""The steps right after the sampling loop with the reference's names and signatures, executing
on the sm_100a kernels of csrc/k_post.cu through the C ABI:

    reconstruct_backbone_partially   /root/reference/AbDock/src/modules/common/geometry.py:450-480
    calc_per_rmsd, calc_avg_rmsd, rank_commoness
                                     /root/reference/AbDock/src/tools/runner/design_for_testset.py:556-589

The reference calls these on CPU tensors (tools/runner/design_for_pdb.py:164-222); CPU inputs are therefore accepted, copied
to the current CUDA device, computed THERE and copied back -- there is no CPU implementation behind these functions.
"""
import torch

from . import _capi

_TABLES = {}


def set_backbone_tables(bb_table, o_table):
    """Ideal backbone geometry: `backbone_atom_coordinates_tensor` (21,3,3) and `bb_oxygen_coordinate_tensor` (21,3) of the
    reference (utils/protein/constants.py:310-320).  `install_into_reference` calls this with the reference's own tables."""
    bb, ox = torch.as_tensor(bb_table, dtype=torch.float32), torch.as_tensor(o_table, dtype=torch.float32)
    if bb.shape != (21, 3, 3) or ox.shape != (21, 3):
        raise ValueError(f'bad table shapes {tuple(bb.shape)} {tuple(ox.shape)}')
    _TABLES.clear()
    _TABLES['cpu'] = (bb.cpu().contiguous(), ox.cpu().contiguous())


def _tables(device, bb_table, o_table):
    if bb_table is not None and o_table is not None:
        return (torch.as_tensor(bb_table, dtype=torch.float32).to(device).contiguous(),
                torch.as_tensor(o_table, dtype=torch.float32).to(device).contiguous())
    if 'cpu' not in _TABLES:
        raise _capi.AboptError('backbone tables not set: call ab_opt_b200.post.set_backbone_tables(bb, o) or '
                               'ab_opt_b200.install_into_reference(...) first, or pass bb_table / o_table')
    key = str(device)
    if key not in _TABLES:
        _TABLES[key] = tuple(x.to(device) for x in _TABLES['cpu'])
    return _TABLES[key]


def _compute_device(t):
    if t.is_cuda:
        return t.device
    if not torch.cuda.is_available():
        raise _capi.AboptError('ab_opt_b200 needs a CUDA device (no CPU fallback)')
    return torch.device('cuda', torch.cuda.current_device())


@torch.no_grad()
def reconstruct_backbone_partially(pos_ctx, R_new, t_new, aa, chain_nb, res_nb, mask_atoms, mask_recons, bb_table=None, o_table=None):
    """pos_ctx (N,L,A,3), R_new (N,L,3,3), t_new (N,L,3), aa / chain_nb / res_nb (N,L), mask_atoms (N,L,A), mask_recons (N,L)
    -> pos_new (N,L,A,3), mask_new (N,L,A) on the device of pos_ctx.  geometry.py:450-480."""
    home = pos_ctx.device
    dev = _compute_device(pos_ctx)
    f32 = lambda x: x.to(dev, torch.float32).contiguous()
    i64 = lambda x: x.to(dev, torch.int64).contiguous()
    u8 = lambda x: x.to(dev, torch.bool).contiguous()
    pos, R, t, aa, ch, rn, ma, mr = f32(pos_ctx), f32(R_new), f32(t_new), i64(aa), i64(chain_nb), i64(res_nb), u8(mask_atoms), u8(mask_recons)
    N, L, A = ma.shape
    if pos.shape != (N, L, A, 3) or R.shape != (N, L, 3, 3) or t.shape != (N, L, 3) or any(x.shape != (N, L) for x in (aa, ch, rn, mr)):
        raise ValueError('bad shapes')
    bb, ox = _tables(dev, bb_table, o_table)
    pos_new, mask_new = torch.empty_like(pos), torch.empty_like(ma)
    with torch.cuda.device(dev):
        _capi.check(_capi.lib().abopt_reconstruct_backbone_partially(
            N, L, A, _capi.ptr(pos), _capi.ptr(R), _capi.ptr(t), _capi.ptr(aa), _capi.ptr(ch), _capi.ptr(rn), _capi.ptr(ma),
            _capi.ptr(mr), _capi.ptr(bb), _capi.ptr(ox), _capi.ptr(pos_new), _capi.ptr(mask_new), _capi.stream_ptr(dev)))
    return pos_new.to(home), mask_new.to(home)


def _structures(structures):
    dev = _compute_device(structures)
    S = structures.to(dev, torch.float32).contiguous()
    if S.dim() != 3 or S.shape[2] != 3:
        raise ValueError(f'structures must be (B, N, 3), got {tuple(S.shape)}')
    return S, dev


@torch.no_grad()
def _rmsd(structures, want_matrix):
    S, dev = _structures(structures)
    B, M, _ = S.shape
    rmsd = torch.empty(B, B, device=dev) if want_matrix else None
    score, avg = torch.empty(B, device=dev), torch.empty((), device=dev)
    with torch.cuda.device(dev):
        _capi.check(_capi.lib().abopt_pairwise_rmsd(B, M, _capi.ptr(S), _capi.ptr(rmsd), _capi.ptr(score), _capi.ptr(avg),
                                                    _capi.stream_ptr(dev)))
    return rmsd, score, avg


def calc_per_rmsd(structures):
    """(B, N, 3) -> (B, B) RMSD of every pair of structures (no superposition).  design_for_testset.py:556-563."""
    return _rmsd(structures, True)[0].to(structures.device)


def calc_avg_rmsd(structures):
    """Mean over the B (B - 1) ordered pairs.  design_for_testset.py:566-570."""
    return _rmsd(structures, False)[2].to(structures.device)


@torch.no_grad()
def rank_commoness(structures, k):
    """Indices of the k structures with the smallest mean RMSD to the others, best first.  design_for_testset.py:573-589."""
    S, dev = _structures(structures)
    B, M, _ = S.shape
    score, rank = torch.empty(B, device=dev), torch.empty(k, device=dev, dtype=torch.int64)
    with torch.cuda.device(dev):
        _capi.check(_capi.lib().abopt_rank_commoness(B, M, _capi.ptr(S), int(k), _capi.ptr(score), _capi.ptr(rank), _capi.stream_ptr(dev)))
    return rank.to(structures.device)
