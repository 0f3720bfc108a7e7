#!/usr/bin/env python
"""Generate the golden fixtures in this directory by RUNNING THE UNMODIFIED REFERENCE
(/root/reference/AbDock/src, imported read-only) on seeded synthetic inputs.

The reference ships no tests or golden vectors, so these fixtures --
outputs of the reference's own code, produced in the build container -- are what pins the
oracle (tests/test_oracle_golden.py) and, through it, the CUDA path (tests/test_gpu_*.py).

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz

Weights are not stored: oracle.weights.make_state_dict(seed, ...) regenerates them
deterministically (numpy RandomState) and is loaded into the reference with strict=True.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import weights, transitions as T   # noqa: E402
from refload import build_reference_fulldpm    # noqa: E402


def npz(name, **arrays):
    out = {}
    for k, v in arrays.items():
        out[k] = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **out)
    print(f'{name}: {os.path.getsize(path) / 1024:.0f} KiB')


@torch.no_grad()
def make_train_forward():
    """FullDPM.forward (training losses, dpm_full.py:156-234) of the unmodified reference, both objectives, at per-complex
    steps t; N=2, L=12 keeps the (N*L, 8191) exponential draw small enough to store."""
    torch.set_num_threads(1)
    seed_w, nl, seed_n = 13, 2, 77
    W = weights.make_state_dict(seed=seed_w, num_layers=nl, flavour='abdock')
    inp = weights.synthetic_inputs(23, 2, 12, gen_slices=((0, 5), (8, 10)), ragged=True)
    t = torch.tensor([57, 3])
    noise = T.draw_step_noise(2, 12, torch.Generator().manual_seed(seed_n))
    keep = {}
    for obj in ('pred_x0', 'pred_noise'):
        model, _ = build_reference_fulldpm(W, num_layers=nl, obj=obj)
        torch.manual_seed(seed_n)      # the reference draws from the global generator
        loss = model(inp['v'], inp['p'], inp['s'], inp['res_feat'], inp['pair_feat'], inp['mask_generate'], inp['mask_res'],
                     denoise_structure=True, denoise_sequence=True, t=t)
        for k, v in loss.items():
            keep[f'{obj}_{k}'] = v
    npz('train_forward.npz', seed_w=seed_w, num_layers=nl, seed_in=23, N=2, L=12, t=t, **keep,
        **{'noise_' + k: v for k, v in noise.items()})


FULL_GRADS = ('eps_net.encoder.blocks.0.spatial_coef', 'eps_net.encoder.blocks.1.spatial_coef', 'eps_net.encoder.blocks.0.proj_pair_bias.weight',
              'eps_net.encoder.blocks.1.layer_norm_2.gamma', 'eps_net.encoder.blocks.0.layer_norm_1.beta', 'eps_net.eps_crd_net.4.weight',
              'eps_net.eps_rot_net.4.weight', 'eps_net.eps_seq_net.4.bias', 'eps_net.current_sequence_embedding.weight',
              'eps_net.prmsd_predictor.linear_3.bias')


def make_train_backward():
    """One training step of the unmodified reference WITH autograd enabled (train.py: loss = sum of the loss dict, weights 1;
    loss.backward()): the grad-enabled losses (log_rotation clamps at -0.999 there, so3.py:12-17), the gradient norm of every
    parameter, the full gradient of a few small ones and d loss / d res_feat, d loss / d pair_feat.  Same inputs and noise as
    train_forward.npz.  This pins oracle.training.loss_and_grads, the target of the CUDA backward."""
    torch.set_num_threads(1)
    seed_w, nl, seed_n = 13, 2, 77
    W = weights.make_state_dict(seed=seed_w, num_layers=nl, flavour='abdock')
    inp = weights.synthetic_inputs(23, 2, 12, gen_slices=((0, 5), (8, 10)), ragged=True)
    t = torch.tensor([57, 3])
    keep = {}
    for obj in ('pred_x0', 'pred_noise'):
        model, _ = build_reference_fulldpm(W, num_layers=nl, obj=obj)
        rf, pf = inp['res_feat'].clone().requires_grad_(True), inp['pair_feat'].clone().requires_grad_(True)
        torch.manual_seed(seed_n)
        loss = model(inp['v'], inp['p'], inp['s'], rf, pf, inp['mask_generate'], inp['mask_res'], denoise_structure=True,
                     denoise_sequence=True, t=t)
        sum(loss.values()).backward()
        for k, v in loss.items():
            keep[f'{obj}_loss_{k}'] = v.detach()
        names = sorted(k for k, _ in model.named_parameters())
        keep[f'{obj}_param_names'] = np.array(names)
        P = dict(model.named_parameters())
        keep[f'{obj}_grad_norms'] = torch.stack([P[k].grad.double().norm() for k in names])
        for k in FULL_GRADS:
            keep[f'{obj}_grad_{k}'] = P[k].grad
        keep[f'{obj}_grad_res_feat'], keep[f'{obj}_grad_pair_feat'] = rf.grad, pf.grad
    npz('train_backward.npz', seed_w=seed_w, num_layers=nl, seed_in=23, N=2, L=12, seed_noise=seed_n, t=t, **keep)


@torch.no_grad()
def make_pair_embed():
    """PairEmbedding.forward (encoders/pair.py:37-101) of the unmodified reference: full-atom (A=15) and backbone+CB (A=5)
    tables, with and without the structure / sequence masks; N=2, L=20, ragged, three chains."""
    torch.set_num_threads(1)
    if os.path.join(os.environ.get('ABOPT_REFERENCE', '/root/reference'), 'AbDock') not in sys.path:
        sys.path.insert(0, os.path.join(os.environ.get('ABOPT_REFERENCE', '/root/reference'), 'AbDock'))
    from src.modules.encoders.pair import PairEmbedding
    from oracle import pair_embed as PE
    seed_w, seed_in, N, L = 3, 5, 2, 20
    inp = PE.synthetic_complex(seed_in, N, L)
    keep = {}
    for A in (15, 5):
        ref = PairEmbedding(64, A)
        ref.load_state_dict(PE.make_state_dict(seed_w, A), strict=True)
        ref.eval()
        args = (inp['aa'], inp['res_nb'], inp['chain_nb'], inp['pos_atoms'], inp['mask_atoms'])
        keep[f'z_a{A}_plain'] = ref(*args)
        keep[f'z_a{A}_masked'] = ref(*args, structure_mask=inp['context_mask'], sequence_mask=inp['context_mask'])
    # ResidueEmbedding.forward (encoders/residue.py:27-94) on the same complexes
    from src.modules.encoders.residue import ResidueEmbedding
    ft = torch.randint(0, 4, (N, L), generator=torch.Generator().manual_seed(seed_in))
    for A in (15, 5):
        ref = ResidueEmbedding(128, A)
        ref.load_state_dict(PE.make_residue_state_dict(seed_w + 1, A), strict=True)
        ref.eval()
        args = (inp['aa'], inp['res_nb'], inp['chain_nb'], inp['pos_atoms'], inp['mask_atoms'], ft)
        keep[f'x_a{A}_plain'] = ref(*args)
        keep[f'x_a{A}_masked'] = ref(*args, structure_mask=inp['context_mask'], sequence_mask=inp['context_mask'])
    # backward of both modules as the reference's autograd computes it (full-atom tables, masked call): d sum(out * G) / d weights
    with torch.enable_grad():
        gz = torch.randn(N, L, L, 64, generator=torch.Generator().manual_seed(31))
        gx = torch.randn(N, L, 128, generator=torch.Generator().manual_seed(32))
        ref = PairEmbedding(64, 15)
        ref.load_state_dict(PE.make_state_dict(seed_w, 15), strict=True)
        (ref(inp['aa'], inp['res_nb'], inp['chain_nb'], inp['pos_atoms'], inp['mask_atoms'], structure_mask=inp['context_mask'],
             sequence_mask=inp['context_mask']) * gz).sum().backward()
        for k, p_ in ref.named_parameters():
            keep['gradz_' + k] = p_.grad
        ref = ResidueEmbedding(128, 15)
        ref.load_state_dict(PE.make_residue_state_dict(seed_w + 1, 15), strict=True)
        (ref(inp['aa'], inp['res_nb'], inp['chain_nb'], inp['pos_atoms'], inp['mask_atoms'], ft, structure_mask=inp['context_mask'],
             sequence_mask=inp['context_mask']) * gx).sum().backward()
        for k, p_ in ref.named_parameters():
            keep['gradx_' + k] = p_.grad
    npz('pair_embed.npz', seed_w=seed_w, seed_in=seed_in, N=N, L=L, fragment_type=ft, **inp, **keep)


@torch.no_grad()
def make_post_loop():
    """The steps after the loop: reconstruct_backbone_partially (geometry.py:450-480) and calc_per_rmsd /
    rank_commoness (tools/runner/design_for_testset.py:556-589) of the unmodified reference.  The ideal backbone tables
    (utils/protein/constants.py:310-320) are stored with the fixture: they are inputs of the C ABI."""
    torch.set_num_threads(1)
    ref_root = os.path.join(os.environ.get('ABOPT_REFERENCE', '/root/reference'), 'AbDock')
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    from src.modules.common.geometry import reconstruct_backbone_partially
    from src.utils.protein import constants as K
    from refload import load_reference_functions
    from oracle import pair_embed as PE, geometry as G
    fn = load_reference_functions('src/tools/runner/design_for_testset.py', ['calc_per_rmsd', 'calc_avg_rmsd', 'rank_commoness'])
    N, L = 2, 20
    inp = PE.synthetic_complex(5, N, L)
    g = torch.Generator().manual_seed(17)
    v = torch.randn(N, L, 3, generator=g)
    t = inp['pos_atoms'][:, :, 1] + torch.randn(N, L, 3, generator=g)
    aa = torch.randint(0, 24, (N, L), generator=g)                  # beyond 20: clamped to UNK by the reference
    rec = ~inp['context_mask']
    pos_new, mask_new = reconstruct_backbone_partially(inp['pos_atoms'], G.so3_exp(v), t, aa, inp['chain_nb'], inp['res_nb'],
                                                       inp['mask_atoms'], rec)
    S = torch.randn(12, 30, 3, generator=g) * 3 + torch.randn(1, 30, 3, generator=g) * 10
    npz('post_loop.npz', bb_table=K.backbone_atom_coordinates_tensor, o_table=K.bb_oxygen_coordinate_tensor, v=v, t=t, aa=aa,
        mask_recons=rec, pos_new=pos_new, mask_new=mask_new, structures=S, rmsd=fn['calc_per_rmsd'](S),
        avg_rmsd=fn['calc_avg_rmsd'](S), rank=fn['rank_commoness'](S, 5))


TRAINED_CKPT = 'AbDock/reproduction/dock_single_cdr/250000.pt'


@torch.no_grad()
def make_trained_slice():
    """A slice of a TRAINED checkpoint of the reference (dock_single_cdr/250000.pt: GABlocks 0-1, the input mixer and the four
    heads of its EpsilonNet, fp32, a few MB) together with what the unmodified reference computes with it: res_feat / pair_feat
    of a synthetic complex through the checkpoint's own residue_embed / pair_embed, then GABlock taps and EpsilonNet.forward of a
    2-layer reference model carrying the slice (parity at trained activation scales, not random-init ones).
    The GPU test loads the same slice into ab_opt_b200.FullDPM (tests/test_gpu_fullsize.py)."""
    torch.set_num_threads(1)
    from test_oracle_vs_reference import _load_ckpt_state
    ref_root = os.environ.get('ABOPT_REFERENCE', '/root/reference')
    if os.path.join(ref_root, 'AbDock') not in sys.path:
        sys.path.insert(0, os.path.join(ref_root, 'AbDock'))
    path = os.path.join(ref_root, TRAINED_CKPT)
    _load_ckpt_state(path)                                  # installs the unpickling stubs
    ck = torch.load(path, map_location='cpu', weights_only=False)['model']
    from src.modules.encoders.pair import PairEmbedding
    from src.modules.encoders.residue import ResidueEmbedding
    from oracle import pair_embed as PE
    N, L, nl, seed_in = 2, 48, 2, 61
    cx = PE.synthetic_complex(seed_in, N, L)
    pe, re_ = PairEmbedding(64, 15), ResidueEmbedding(128, 15)
    pe.load_state_dict({k[len('pair_embed.'):]: v for k, v in ck.items() if k.startswith('pair_embed.')}, strict=True)
    re_.load_state_dict({k[len('residue_embed.'):]: v for k, v in ck.items() if k.startswith('residue_embed.')}, strict=True)
    ft = torch.randint(1, 4, (N, L), generator=torch.Generator().manual_seed(seed_in))
    a = (cx['aa'], cx['res_nb'], cx['chain_nb'], cx['pos_atoms'], cx['mask_atoms'])
    ctx = cx['context_mask']
    res_feat = re_.eval()(*a, ft, structure_mask=ctx, sequence_mask=ctx)
    pair_feat = pe.eval()(*a, structure_mask=ctx, sequence_mask=ctx)
    mask_res = cx['mask_atoms'][:, :, 1]
    mask_gen = (~ctx) & mask_res
    # the slice: everything of `diffusion.eps_net` except blocks >= nl (+ the diffusion buffers, regenerated by the oracle)
    W = weights.make_state_dict(seed=0, num_layers=nl, flavour='abdock')
    sl = {}
    for k in W:
        if k.startswith('eps_net.'):
            sl[k] = ck['diffusion.' + k].float().contiguous()
            assert sl[k].shape == W[k].shape, k
    W.update(sl)
    model, m = build_reference_fulldpm(W, num_layers=nl, obj='pred_x0')
    g = torch.Generator().manual_seed(7)
    from oracle.geometry import uniform_so3_from_gauss4
    v = uniform_so3_from_gauss4(torch.randn(N, L, 4, generator=g))
    p = cx['pos_atoms'][:, :, 1] - cx['pos_atoms'][:, :, 1][mask_res].mean(0)          # CA, roughly centred, Angstrom
    s = cx['aa'].clamp(max=19)
    s = torch.where(mask_res, s, torch.full_like(s, 21))
    R, tpos = m['so3'].so3vec_to_rotation(v), p / 10.0
    blk = model.eps_net.encoder.blocks[0]
    logits = blk._node_logits(res_feat) + blk._pair_logits(pair_feat) + blk._spatial_logits(R, tpos, res_feat)
    alpha = m['ga']._alpha_from_logits(logits * np.sqrt(1 / 3), mask_res)
    feat = torch.cat([blk._pair_aggregation(alpha, pair_feat), blk._node_aggregation(alpha, res_feat),
                      blk._spatial_aggregation(alpha, R, tpos, res_feat)], -1)
    tstep = 35
    beta = W['trans_pos.var_sched.betas'][tstep].expand(N)
    out = model.eps_net(v, tpos, s, res_feat, pair_feat, beta, mask_gen, mask_res)
    npz('trained_slice.npz', N=N, L=L, num_layers=nl, t=tstep, ckpt=TRAINED_CKPT, v=v, p=p, s=s, res_feat=res_feat, pair_feat=pair_feat,
        mask_generate=mask_gen, mask_res=mask_res, alpha=alpha, feat=feat, x_out=blk(R, tpos, res_feat, pair_feat, mask_res),
        enc_out=model.eps_net.encoder(R, tpos, res_feat, pair_feat, mask_res), v_next=out[0], R_next=out[1], eps_pos=out[2],
        c_denoised=out[3], prmsd_logits=out[4], **{'W.' + k: v_ for k, v_ in sl.items()})


@torch.no_grad()
def main():
    torch.set_num_threads(1)          # single-thread reference: reproducible reduction order
    # ---------------------------------------------------------------- GABlock / GAEncoder
    seed_w, nl = 11, 2
    W = weights.make_state_dict(seed=seed_w, num_layers=nl, flavour='abdock')
    model, m = build_reference_fulldpm(W, num_layers=nl, obj='pred_x0')
    inp = weights.synthetic_inputs(21, 2, 24, gen_slices=((8, 14),), ragged=True)
    R = m['so3'].so3vec_to_rotation(inp['v'])
    t = inp['p'] / 10.0
    blk = model.eps_net.encoder.blocks[0]
    x, z, mask = inp['res_feat'], inp['pair_feat'], inp['mask_res']
    logits = blk._node_logits(x) + blk._pair_logits(z) + blk._spatial_logits(R, t, x)
    alpha = m['ga']._alpha_from_logits(logits * np.sqrt(1 / 3), mask)
    feat = torch.cat([blk._pair_aggregation(alpha, z), blk._node_aggregation(alpha, x),
                      blk._spatial_aggregation(alpha, R, t, x)], -1)
    npz('ga_block.npz', seed_w=seed_w, num_layers=nl, seed_in=21, N=2, L=24,
        logits=logits, alpha=alpha, feat=feat, x_out=blk(R, t, x, z, mask),
        enc_out=model.eps_net.encoder(R, t, x, z, mask))

    # ---------------------------------------------------------------- EpsilonNet (AbDock flavour)
    beta = W['trans_pos.var_sched.betas'][57].expand(2)
    out = model.eps_net(inp['v'], t, inp['s'], x, z, beta, inp['mask_generate'], mask)
    npz('eps_net_abdock.npz', seed_w=seed_w, num_layers=nl, seed_in=21, N=2, L=24, t=57,
        v_next=out[0], R_next=out[1], eps_pos=out[2], c_denoised=out[3], prmsd_logits=out[4],
        prmsd=model.prmsd.compute_prmsd(out[4]))

    # ---------------------------------------------------------------- transitions, replayed noise
    # One reverse step's three denoise calls at t = 57 (histogram branch), t = 3 (Gaussian
    # branch: sigma <= 0.1) and t = 1 (noise switched off), N=1, L=12 so that the
    # (N*L, 8191) exponential draw stays small enough to store.
    sm = weights.synthetic_inputs(22, 1, 12, gen_slices=((3, 9),))
    dpm_ref = m['dpm_full']
    for tstep in (57, 3, 1):
        seed_n = 1000 + tstep
        noise = T.draw_step_noise(1, 12, torch.Generator().manual_seed(seed_n))
        tt = torch.full((1,), tstep, dtype=torch.long)
        v_t, p_t, s_t = sm['v'], sm['p'] / 10.0, sm['s']
        v_net = m['so3'].rotation_to_so3vec(m['so3'].so3vec_to_rotation(sm['v'] * 0.7))
        p_pred = p_t * 0.9 + 0.05
        c0 = torch.softmax(sm['res_feat'][..., :20], -1)
        torch.manual_seed(seed_n)      # the reference draws from the global generator
        eps_p = model.trans_pos.pred_noise_from_start(p_t, p_pred, sm['mask_generate'], tt)
        v_next = model.trans_rot.denoise(v_t, v_net, sm['mask_generate'], tt)
        p_next = model.trans_pos.denoise(p_t, eps_p, sm['mask_generate'], tt)
        post, s_next = model.trans_seq.denoise(s_t, c0, sm['mask_generate'], tt)
        ppl = dpm_ref.calc_perplexity(post, sm['mask_generate'])
        npz(f'transitions_t{tstep}.npz', seed_in=22, N=1, L=12, t=tstep, v_net=v_net, p_pred=p_pred,
            c0=c0, eps_p=eps_p, v_next=v_next, p_next=p_next, post=post, s_next=s_next, ppl=ppl,
            **{'noise_' + k: v for k, v in noise.items()})

    # ---------------------------------------------------------------- forward noising (optimize path)
    seed_n = 2024
    tt = torch.full((1,), 40, dtype=torch.long)
    noise = T.draw_step_noise(1, 12, torch.Generator().manual_seed(seed_n))
    torch.manual_seed(seed_n)
    v_noisy, _ = model.trans_rot.add_noise(sm['v'], sm['mask_generate'], tt)
    p_noisy, _ = model.trans_pos.add_noise(sm['p'] / 10.0, sm['mask_generate'], tt)
    _, s_noisy = model.trans_seq.add_noise(sm['s'], sm['mask_generate'], tt)
    npz('add_noise_t40.npz', seed_in=22, N=1, L=12, t=40, v_noisy=v_noisy, p_noisy=p_noisy,
        s_noisy=s_noisy, **{'noise_' + k: v for k, v in noise.items()})

    # ---------------------------------------------------------------- FullDPM.sample, first steps
    # Same-seed run of the reference; we keep the initial state and the first two reverse
    # steps (later steps decorrelate chaotically even reference-vs-reference, SURVEY finding 4)
    # plus the whole amino-acid trajectory of the generated residues.
    seed_s = 4242
    torch.manual_seed(seed_s)
    traj = model.sample(inp['v'], inp['p'], inp['s'], x, z, inp['mask_generate'], mask)
    keep = {}
    for k in (100, 99, 98):
        keep[f'v_{k}'], keep[f'p_{k}'], keep[f's_{k}'] = traj[k][0], traj[k][1], traj[k][2]
    keep['prmsd_99'], keep['ppl_99'] = traj[99][3], traj[99][4]
    npz('sample_first_steps.npz', seed_w=seed_w, num_layers=nl, seed_in=21, N=2, L=24, seed_s=seed_s,
        **keep)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'train_forward':      # only the training fixture (the others are unchanged)
        make_train_forward()
    elif len(sys.argv) > 1 and sys.argv[1] == 'pair_embed':       # only the pair-featurisation fixture
        make_pair_embed()
    elif len(sys.argv) > 1 and sys.argv[1] == 'post_loop':
        make_post_loop()
    elif len(sys.argv) > 1 and sys.argv[1] == 'train_backward':
        make_train_backward()
    elif len(sys.argv) > 1 and sys.argv[1] == 'trained_slice':
        make_trained_slice()
    else:
        main()
        make_train_forward()
        make_pair_embed()
        make_post_loop()
        make_train_backward()
        make_trained_slice()
