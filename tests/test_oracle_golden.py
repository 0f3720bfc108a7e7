"""Pins the CPU oracle against outputs of the UNMODIFIED reference (tests/golden/*.npz,
produced by tests/golden/make_golden.py).  Runs anywhere (no GPU, no /root/reference)."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import weights, ipa, epsnet, sampler, transitions as T, geometry as G


def load(golden_dir, name):
    d = np.load(os.path.join(golden_dir, name))
    return {k: torch.from_numpy(d[k]) if d[k].ndim else d[k].item() for k in d.files}


@pytest.fixture(scope='module')
def case(golden_dir):
    g = load(golden_dir, 'ga_block.npz')
    W = weights.make_state_dict(seed=g['seed_w'], num_layers=g['num_layers'], flavour='abdock')
    inp = weights.synthetic_inputs(g['seed_in'], g['N'], g['L'], gen_slices=((8, 14),), ragged=True)
    return W, inp


@pytest.mark.parametrize('materialize', [True, False])
def test_ga_block_matches_reference(golden_dir, case, materialize):
    g = load(golden_dir, 'ga_block.npz')
    W, inp = case
    R, t = G.so3_exp(inp['v']), inp['p'] / 10.0
    out, parts = ipa.ga_block(W, 'eps_net.encoder.blocks.0.', R, t, inp['res_feat'], inp['pair_feat'],
                              inp['mask_res'], materialize=materialize, return_parts=True)
    torch.testing.assert_close(parts['logits'], g['logits'], rtol=1e-5, atol=2e-5)
    torch.testing.assert_close(parts['alpha'], g['alpha'], rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(parts['feat'], g['feat'], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(out, g['x_out'], rtol=1e-4, atol=1e-5)
    enc = ipa.ga_encoder(W, 'eps_net.encoder.', R, t, inp['res_feat'], inp['pair_feat'], inp['mask_res'],
                         g['num_layers'], materialize)
    torch.testing.assert_close(enc, g['enc_out'], rtol=1e-4, atol=2e-5)


def test_masked_rows_have_zero_attention(case):
    W, inp = case
    R, t = G.so3_exp(inp['v']), inp['p'] / 10.0
    _, parts = ipa.ga_block(W, 'eps_net.encoder.blocks.0.', R, t, inp['res_feat'], inp['pair_feat'],
                            inp['mask_res'], return_parts=True)
    a = parts['alpha']
    assert (~inp['mask_res']).any()
    assert a[~inp['mask_res']].abs().max() == 0
    assert a.transpose(1, 2)[~inp['mask_res']].abs().max() == 0          # masked keys get no weight
    torch.testing.assert_close(a[inp['mask_res']].sum(1), torch.ones_like(a[inp['mask_res']].sum(1)))


def test_eps_net_matches_reference(golden_dir, case):
    g = load(golden_dir, 'eps_net_abdock.npz')
    W, inp = case
    beta = W['trans_pos.var_sched.betas'][g['t']].expand(g['N'])
    out = epsnet.eps_net(W, inp['v'], inp['p'] / 10.0, inp['s'], inp['res_feat'], inp['pair_feat'], beta,
                         inp['mask_generate'], inp['mask_res'])
    torch.testing.assert_close(out[1], g['R_next'], rtol=0, atol=2e-6)
    torch.testing.assert_close(out[2], g['eps_pos'], rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(out[3], g['c_denoised'], rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(out[4], g['prmsd_logits'], rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(epsnet.prmsd_score(out[4]), g['prmsd'], rtol=1e-5, atol=1e-5)
    # v_next goes through the ill-conditioned log map: compare as rotations
    torch.testing.assert_close(G.so3_exp(out[0]), G.so3_exp(g['v_next']), rtol=0, atol=1e-5)
    # context residues keep their orientation bit-for-bit
    keep = ~inp['mask_generate']
    assert torch.equal(out[0][keep], inp['v'][keep])


@pytest.mark.parametrize('tstep', [57, 3, 1])
def test_transitions_replayed_noise(golden_dir, tstep):
    g = load(golden_dir, f'transitions_t{tstep}.npz')
    W = T.diffusion_buffers(100)
    sm = weights.synthetic_inputs(g['seed_in'], g['N'], g['L'], gen_slices=((3, 9),))
    noise = {k[6:]: v for k, v in g.items() if k.startswith('noise_')}
    tt = torch.full((g['N'],), tstep, dtype=torch.long)
    v_t, p_t, s_t, mg = sm['v'], sm['p'] / 10.0, sm['s'], sm['mask_generate']
    eps_p = T.pos_pred_noise_from_start(W, p_t, g['p_pred'], mg, tt)
    torch.testing.assert_close(eps_p, g['eps_p'], rtol=1e-6, atol=1e-6)
    v_next = T.rot_denoise(W, v_t, g['v_net'], mg, tt, noise)
    torch.testing.assert_close(G.so3_exp(v_next), G.so3_exp(g['v_next']), rtol=0, atol=1e-5)
    assert torch.equal(v_next[~mg], v_t[~mg])
    p_next = T.pos_denoise(W, p_t, g['eps_p'], mg, tt, noise['z_pos'])
    torch.testing.assert_close(p_next, g['p_next'], rtol=1e-6, atol=1e-6)
    post, s_next = T.seq_denoise(W, s_t, g['c0'], mg, tt, noise['expo_seq'])
    torch.testing.assert_close(post, g['post'], rtol=1e-6, atol=1e-8)
    assert torch.equal(s_next, g['s_next'])                               # bit-exact indices
    torch.testing.assert_close(epsnet.perplexity(post, mg), g['ppl'], rtol=1e-6, atol=1e-7)
    if tstep == 1:      # last step adds no noise: pure network update
        torch.testing.assert_close(G.so3_exp(v_next[mg]), G.so3_exp(g['v_net'][mg]), rtol=0, atol=1e-5)


def test_add_noise_replayed(golden_dir):
    g = load(golden_dir, 'add_noise_t40.npz')
    W = T.diffusion_buffers(100)
    sm = weights.synthetic_inputs(g['seed_in'], g['N'], g['L'], gen_slices=((3, 9),))
    noise = {k[6:]: v for k, v in g.items() if k.startswith('noise_')}
    tt = torch.full((g['N'],), g['t'], dtype=torch.long)
    mg = sm['mask_generate']
    v_noisy, _ = T.rot_add_noise(W, sm['v'], mg, tt, noise)
    torch.testing.assert_close(G.so3_exp(v_noisy), G.so3_exp(g['v_noisy']), rtol=0, atol=1e-5)
    torch.testing.assert_close(T.pos_add_noise(W, sm['p'] / 10.0, mg, tt, noise['z_pos']), g['p_noisy'],
                               rtol=1e-6, atol=1e-6)
    _, s_noisy = T.seq_add_noise(W, sm['s'], mg, tt, noise['expo_seq'])
    assert torch.equal(s_noisy, g['s_noisy'])


def test_multinomial_emulation_matches_torch():
    """argmax(p / Exp(1)) IS torch.multinomial(p, 1) for the same generator state."""
    p = torch.rand(64, 20) + 1e-3
    a = torch.multinomial(p, 1, generator=torch.Generator().manual_seed(5)).squeeze(-1)
    q = torch.empty(64, 20).exponential_(1, generator=torch.Generator().manual_seed(5))
    assert torch.equal(a, T.multinomial_from_exp(p, q))
    p = torch.rand(8, 8192)[:, :-1]                                       # non-contiguous slice, as so3.py:123
    a = torch.multinomial(p, 1, generator=torch.Generator().manual_seed(6)).squeeze(-1)
    q = torch.empty(8, 8191).exponential_(1, generator=torch.Generator().manual_seed(6))
    assert torch.equal(a, T.multinomial_from_exp(p, q))


def test_sample_same_seed_first_steps(golden_dir, case):
    """Same seed -> same draws -> the oracle reproduces the reference's trajectory start.
    (Relies on torch's CPU generator giving the same stream as when the fixture was made.)"""
    g = load(golden_dir, 'sample_first_steps.npz')
    W, inp = case
    gen = torch.Generator().manual_seed(g['seed_s'])
    traj = sampler.sample(W, inp['v'], inp['p'], inp['s'], inp['res_feat'], inp['pair_feat'],
                          inp['mask_generate'], inp['mask_res'], obj='pred_x0', gen=gen, stop_at=98)
    assert torch.equal(traj[100][2], g['s_100'])
    torch.testing.assert_close(traj[100][1], g['p_100'], rtol=1e-6, atol=1e-5)
    torch.testing.assert_close(G.so3_exp(traj[100][0]), G.so3_exp(g['v_100']), rtol=0, atol=1e-5)
    for k in (99, 98):
        assert torch.equal(traj[k][2], g[f's_{k}'])
        torch.testing.assert_close(traj[k][1], g[f'p_{k}'], rtol=1e-4, atol=1e-3)   # Angstrom
    torch.testing.assert_close(traj[99][3], g['prmsd_99'], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(traj[99][4], g['ppl_99'], rtol=1e-5, atol=1e-6)


def test_schedule_properties():
    s = T.variance_schedule(100)
    assert s['betas'][0] == 0 and (s['betas'][1:] > 0).all() and s['betas'].max() <= 0.999
    assert (s['alpha_bars'][1:] < s['alpha_bars'][:-1]).all()
    assert math.isclose(s['alpha_bars'][0].item(), 1.0)


@pytest.mark.parametrize('obj', ['pred_x0', 'pred_noise'])
def test_training_forward_matches_reference(golden_dir, obj):
    """oracle.training.loss_forward vs FullDPM.forward of the unmodified reference (tests/golden/train_forward.npz)."""
    from oracle import training
    g = load(golden_dir, 'train_forward.npz')
    W = weights.make_state_dict(seed=g['seed_w'], num_layers=g['num_layers'], flavour='abdock')
    inp = weights.synthetic_inputs(g['seed_in'], g['N'], g['L'], gen_slices=((0, 5), (8, 10)), ragged=True)
    noise = {k[len('noise_'):]: v for k, v in g.items() if k.startswith('noise_')}
    got = training.loss_forward(W, inp['v'], inp['p'], inp['s'], inp['res_feat'], inp['pair_feat'], inp['mask_generate'],
                                inp['mask_res'], True, True, g['t'], noise, flavour='abdock', obj=obj)
    want = {k[len(obj) + 1:]: v for k, v in g.items() if k.startswith(obj + '_')}
    assert sorted(got) == sorted(want)
    for k in want:
        torch.testing.assert_close(got[k], torch.as_tensor(want[k]), rtol=2e-5, atol=2e-6, msg=lambda m, k=k: f"{k}: {m}")


# ------------------------------------------------------------------------------------------ pair featurisation
@pytest.mark.parametrize('A', [15, 5])
@pytest.mark.parametrize('masked', [False, True])
def test_pair_embedding_matches_reference(golden_dir, A, masked):
    """oracle.pair_embed.pair_embedding vs PairEmbedding.forward of the unmodified reference (tests/golden/pair_embed.npz)."""
    from oracle import pair_embed as PE
    g = load(golden_dir, 'pair_embed.npz')
    W = PE.make_state_dict(g['seed_w'], A)
    inp = PE.synthetic_complex(g['seed_in'], g['N'], g['L'])
    for k in ('aa', 'res_nb', 'chain_nb', 'pos_atoms', 'mask_atoms', 'context_mask'):       # the generator is part of the pin
        assert torch.equal(inp[k], g[k]), k
    m = inp['context_mask'] if masked else None
    z = PE.pair_embedding(W, inp['aa'], inp['res_nb'], inp['chain_nb'], inp['pos_atoms'], inp['mask_atoms'], m, m)
    ref = g[f'z_a{A}_' + ('masked' if masked else 'plain')]
    off = ~torch.eye(g['L'], dtype=torch.bool)[None, :, :, None]
    torch.testing.assert_close(z * off, ref * off, rtol=1e-5, atol=2e-6)
    # i == j: the sign of the inter-residue dihedral is the sign of a triple product that is zero in exact arithmetic
    # (geometry.py:268 with p0 == p3), i.e. rounding noise in the reference itself; +-0.0014 rad moves z by < 1e-3
    torch.testing.assert_close(z, ref, rtol=0, atol=2e-3)
    assert (z[~inp['mask_atoms'][:, :, 1]] == 0).all()


@pytest.mark.parametrize('A', [15, 5])
@pytest.mark.parametrize('masked', [False, True])
def test_residue_embedding_matches_reference(golden_dir, A, masked):
    """oracle.pair_embed.residue_embedding vs ResidueEmbedding.forward of the unmodified reference (same fixture file)."""
    from oracle import pair_embed as PE
    g = load(golden_dir, 'pair_embed.npz')
    W = PE.make_residue_state_dict(g['seed_w'] + 1, A)
    inp = PE.synthetic_complex(g['seed_in'], g['N'], g['L'])
    m = inp['context_mask'] if masked else None
    x = PE.residue_embedding(W, inp['aa'], inp['res_nb'], inp['chain_nb'], inp['pos_atoms'], inp['mask_atoms'], g['fragment_type'], m, m)
    torch.testing.assert_close(x, g[f'x_a{A}_' + ('masked' if masked else 'plain')], rtol=1e-5, atol=1e-6)
    assert (x[~inp['mask_atoms'][:, :, 1]] == 0).all()


# ------------------------------------------------------------------------------------------ after the loop
def test_post_loop_matches_reference(golden_dir):
    """oracle.post vs reconstruct_backbone_partially / calc_per_rmsd / calc_avg_rmsd / rank_commoness of the reference."""
    from oracle import post, pair_embed as PE
    g = load(golden_dir, 'post_loop.npz')
    inp = PE.synthetic_complex(5, 2, 20)
    pos_new, mask_new = post.reconstruct_backbone_partially(inp['pos_atoms'], G.so3_exp(g['v']), g['t'], g['aa'], inp['chain_nb'],
                                                            inp['res_nb'], inp['mask_atoms'], g['mask_recons'], g['bb_table'], g['o_table'])
    torch.testing.assert_close(pos_new, g['pos_new'], rtol=1e-6, atol=1e-5)
    assert torch.equal(mask_new, g['mask_new'])
    torch.testing.assert_close(post.pairwise_rmsd(g['structures']), g['rmsd'], rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(post.average_rmsd(g['structures']), torch.tensor(g['avg_rmsd']), rtol=1e-6, atol=0)
    assert torch.equal(post.rank_commonness(g['structures'], 5), g['rank'])


# ------------------------------------------------------------------------------------------ training step with autograd
@pytest.mark.parametrize('obj', ['pred_x0', 'pred_noise'])
def test_training_backward_matches_reference(golden_dir, obj):
    """oracle.training.loss_and_grads vs one training step of the unmodified reference with autograd ENABLED
    (tests/golden/train_backward.npz): the grad-enabled losses (log_rotation clamps at -0.999 there, so3.py:12-17 -- they differ
    from the no_grad losses of train_forward.npz by up to 2 %), the gradient norm of all 71 parameters, the full gradient of ten
    of them and the gradients with respect to res_feat / pair_feat."""
    from oracle import training
    d = np.load(os.path.join(golden_dir, 'train_backward.npz'))
    g = {k: (torch.from_numpy(d[k]) if d[k].dtype.kind != 'U' and d[k].ndim else d[k]) for k in d.files}
    W = weights.make_state_dict(seed=int(g['seed_w']), num_layers=int(g['num_layers']), flavour='abdock')
    inp = weights.synthetic_inputs(int(g['seed_in']), int(g['N']), int(g['L']), gen_slices=((0, 5), (8, 10)), ragged=True)
    noise = T.draw_step_noise(int(g['N']), int(g['L']), torch.Generator().manual_seed(int(g['seed_noise'])))
    loss, grads, g_res, g_pair = training.loss_and_grads(W, inp['v'], inp['p'], inp['s'], inp['res_feat'], inp['pair_feat'],
                                                         inp['mask_generate'], inp['mask_res'], True, True, g['t'], noise,
                                                         flavour='abdock', obj=obj)
    for k, v in loss.items():
        torch.testing.assert_close(v, torch.as_tensor(g[f'{obj}_loss_{k}'].item()), rtol=2e-5, atol=2e-6, msg=lambda m, k=k: f'{k}: {m}')
    names = [str(x) for x in g[f'{obj}_param_names']]
    assert sorted(grads) == names
    norms = torch.stack([grads[k].double().norm() for k in names])
    torch.testing.assert_close(norms, g[f'{obj}_grad_norms'], rtol=1e-4, atol=1e-9)
    for key in d.files:
        if key.startswith(f'{obj}_grad_eps_net.'):
            want = g[key]
            got = grads[key[len(obj) + 6:]]
            assert (got - want).abs().max() <= 2e-5 * want.abs().max() + 1e-9, key
    for got, want in ((g_res, g[f'{obj}_grad_res_feat']), (g_pair, g[f'{obj}_grad_pair_feat'])):
        assert (got - want).abs().max() <= 2e-5 * want.abs().max()
    # and the no_grad losses are NOT these: the clamp matters
    plain = training.loss_forward(W, inp['v'], inp['p'], inp['s'], inp['res_feat'], inp['pair_feat'], inp['mask_generate'],
                                  inp['mask_res'], True, True, g['t'], noise, flavour='abdock', obj=obj)
    assert abs(float(plain['rot']) - float(loss['rot'])) > 1e-3


@pytest.mark.parametrize('dtype,tol', [(torch.float64, 1e-9), (torch.float32, 2e-5)])
def test_hand_written_block_backward_matches_autograd(dtype, tol):
    """oracle.ipa_backward.ga_block_backward (the explicit formulas the CUDA backward will implement) vs torch autograd through
    oracle.ipa.ga_block (pinned to the reference's GABlock): d x, d z and all 21 weight gradients of a block; ragged mask."""
    from oracle import ipa_backward
    W = weights.cast(weights.make_state_dict(seed=5, num_layers=1, flavour='abdesign'), dtype)
    inp = weights.synthetic_inputs(9, 2, 20, gen_slices=((4, 9),), ragged=True, dtype=dtype)
    prefix = 'eps_net.encoder.blocks.0.'
    R, t = G.so3_exp(inp['v']), inp['p'] / 10
    keys = [k for k in W if k.startswith(prefix)]
    Wg = dict(W)
    for k in keys:
        Wg[k] = W[k].clone().requires_grad_(True)
    x, z = inp['res_feat'].clone().requires_grad_(True), inp['pair_feat'].clone().requires_grad_(True)
    out = ipa.ga_block(Wg, prefix, R, t, x, z, inp['mask_res'], materialize=False)
    g_out = torch.randn(out.shape, dtype=dtype, generator=torch.Generator().manual_seed(1))
    out.backward(g_out)
    g_x, g_z, g_w = ipa_backward.ga_block_backward(W, prefix, R, t, inp['res_feat'], inp['pair_feat'], inp['mask_res'], g_out)
    assert sorted(g_w) == sorted(keys)
    for name, got, want in [('x', g_x, x.grad), ('z', g_z, z.grad)] + [(k, g_w[k], Wg[k].grad) for k in keys]:
        assert (got - want).abs().max() <= tol * want.abs().max() + 1e-30, name
    assert (g_z[~inp['mask_res']] == 0).all()                    # padded query rows receive no gradient


@pytest.mark.parametrize('dtype,tol', [(torch.float64, 1e-9), (torch.float32, 2e-5)])
def test_hand_written_training_step_matches_autograd(dtype, tol):
    """oracle.epsnet_backward.training_step_abdesign -- losses, heads, quaternion update, encoder, mixer and embedding gradients
    written out by hand -- vs oracle.training.loss_and_grads (= the reference's autograd step): the three losses, all 63
    parameter gradients and d / d res_feat, d / d pair_feat of one AbDesign-flavour training step."""
    from oracle import training, epsnet_backward
    W = weights.cast(weights.make_state_dict(seed=13, num_layers=2, flavour='abdesign'), dtype)
    inp = weights.synthetic_inputs(23, 2, 12, gen_slices=((0, 5), (8, 10)), ragged=True, dtype=dtype)
    t = torch.tensor([57, 3])
    noise = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in T.draw_step_noise(2, 12, torch.Generator().manual_seed(77)).items()}
    a = (W, inp['v'], inp['p'], inp['s'], inp['res_feat'], inp['pair_feat'], inp['mask_generate'], inp['mask_res'])
    loss, grads, g_res, g_pair = training.loss_and_grads(*a, True, True, t, noise, flavour='abdesign', obj='pred_noise')
    loss2, grads2, g_res2, g_pair2 = epsnet_backward.training_step_abdesign(*a, t, noise)
    assert sorted(loss) == sorted(loss2) and sorted(grads) == sorted(grads2)
    for k in loss:
        torch.testing.assert_close(loss2[k], loss[k], rtol=10 * tol, atol=0)
    for name, got, want in [('res_feat', g_res2, g_res), ('pair_feat', g_pair2, g_pair)] + [(k, grads2[k], grads[k]) for k in grads]:
        assert (got - want).abs().max() <= tol * want.abs().max() + 1e-30, name


@pytest.mark.parametrize('obj', ['pred_x0', 'pred_noise'])
def test_hand_written_training_step_matches_the_reference_gradients(golden_dir, obj):
    """oracle.epsnet_backward.training_step (AbDock flavour: + pRMSD head and loss, distance loss; no autograd anywhere) vs the
    gradients the UNMODIFIED REFERENCE computed with loss.backward() (tests/golden/train_backward.npz): losses, the gradient norm
    of all 71 parameters, ten full gradients, d / d res_feat and d / d pair_feat."""
    from oracle import epsnet_backward
    d = np.load(os.path.join(golden_dir, 'train_backward.npz'))
    g = {k: (torch.from_numpy(d[k]) if d[k].dtype.kind != 'U' and d[k].ndim else d[k]) for k in d.files}
    W = weights.make_state_dict(seed=int(g['seed_w']), num_layers=int(g['num_layers']), flavour='abdock')
    inp = weights.synthetic_inputs(int(g['seed_in']), int(g['N']), int(g['L']), gen_slices=((0, 5), (8, 10)), ragged=True)
    noise = T.draw_step_noise(int(g['N']), int(g['L']), torch.Generator().manual_seed(int(g['seed_noise'])))
    loss, grads, g_res, g_pair = epsnet_backward.training_step(W, inp['v'], inp['p'], inp['s'], inp['res_feat'], inp['pair_feat'],
                                                               inp['mask_generate'], inp['mask_res'], g['t'], noise, flavour='abdock', obj=obj)
    for k, v in loss.items():
        torch.testing.assert_close(v, torch.as_tensor(g[f'{obj}_loss_{k}'].item()), rtol=2e-5, atol=2e-6, msg=lambda m, k=k: f'{k}: {m}')
    names = [str(x) for x in g[f'{obj}_param_names']]
    assert sorted(grads) == names
    torch.testing.assert_close(torch.stack([grads[k].double().norm() for k in names]), g[f'{obj}_grad_norms'], rtol=1e-4, atol=1e-9)
    for key in d.files:
        if key.startswith(f'{obj}_grad_eps_net.'):
            want, got = g[key], grads[key[len(obj) + 6:]]
            assert (got - want).abs().max() <= 2e-5 * want.abs().max() + 1e-9, key
    for got, want in ((g_res, g[f'{obj}_grad_res_feat']), (g_pair, g[f'{obj}_grad_pair_feat'])):
        assert (got - want).abs().max() <= 2e-5 * want.abs().max()


def test_embedding_backward_matches_reference(golden_dir):
    """oracle.pair_embed.embedding_grads vs the reference's own autograd through PairEmbedding / ResidueEmbedding (weight
    gradients of sum(out * G) stored in tests/golden/pair_embed.npz): the featurisation's share of a training step's backward."""
    from oracle import pair_embed as PE
    g = load(golden_dir, 'pair_embed.npz')
    inp = PE.synthetic_complex(g['seed_in'], g['N'], g['L'])
    m = inp['context_mask']
    base = (inp['aa'], inp['res_nb'], inp['chain_nb'], inp['pos_atoms'], inp['mask_atoms'])
    gz = torch.randn(g['N'], g['L'], g['L'], 64, generator=torch.Generator().manual_seed(31))
    gx = torch.randn(g['N'], g['L'], 128, generator=torch.Generator().manual_seed(32))
    got = PE.embedding_grads('pair', PE.make_state_dict(g['seed_w'], 15), gz, *base, m, m)
    want = {k[len('gradz_'):]: v for k, v in g.items() if k.startswith('gradz_')}
    assert sorted(got) == sorted(want) and len(want) == 13
    for k in want:
        assert (got[k] - want[k]).abs().max() <= 2e-5 * want[k].abs().max() + 1e-12, k
    got = PE.embedding_grads('residue', PE.make_residue_state_dict(g['seed_w'] + 1, 15), gx, *base, g['fragment_type'], m, m)
    want = {k[len('gradx_'):]: v for k, v in g.items() if k.startswith('gradx_')}
    assert sorted(got) == sorted(want) and len(want) == 10
    for k in want:
        assert (got[k] - want[k]).abs().max() <= 2e-5 * want[k].abs().max() + 1e-12, k


def trained_slice(golden_dir):
    """The trained-checkpoint slice fixture -> (state dict of a 2-layer AbDock-flavour FullDPM, tensors)."""
    g = load(golden_dir, 'trained_slice.npz')
    W = weights.make_state_dict(seed=0, num_layers=g['num_layers'], flavour='abdock')
    W.update({k[2:]: v for k, v in g.items() if k.startswith('W.')})
    return W, g


def test_trained_checkpoint_slice_matches_reference(golden_dir):
    """The oracle with TRAINED weights (blocks 0-1, mixer and heads of dock_single_cdr/250000.pt) on res_feat / pair_feat produced
    by the checkpoint's own embeddings, against what the unmodified reference computed."""
    W, g = trained_slice(golden_dir)
    R, t = G.so3_exp(g['v']), g['p'] / 10.0
    out, parts = ipa.ga_block(W, 'eps_net.encoder.blocks.0.', R, t, g['res_feat'], g['pair_feat'], g['mask_res'],
                              materialize=False, return_parts=True)
    torch.testing.assert_close(parts['alpha'], g['alpha'], rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(parts['feat'], g['feat'], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(out, g['x_out'], rtol=1e-4, atol=2e-5)
    beta = W['trans_pos.var_sched.betas'][g['t']].expand(g['N'])
    o = epsnet.eps_net(W, g['v'], t, g['s'], g['res_feat'], g['pair_feat'], beta, g['mask_generate'], g['mask_res'], materialize=False)
    torch.testing.assert_close(o[1], g['R_next'], rtol=0, atol=2e-5)
    torch.testing.assert_close(o[2], g['eps_pos'], rtol=1e-4, atol=5e-6)
    torch.testing.assert_close(o[3], g['c_denoised'], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(o[4], g['prmsd_logits'], rtol=1e-4, atol=2e-5)
