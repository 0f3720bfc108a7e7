"""GPU parity of the hand-written backward pass (csrc/k_backward.cu): one GABlock, the whole
training step of both flavours against the oracle's hand-written step (pinned to the reference's autograd), the gradients the
UNMODIFIED REFERENCE computed with loss.backward() (tests/golden/train_backward.npz), and the autograd integration
(`loss.backward()` on the dict FullDPM.forward returns).  Run on the B200 box:  pytest tests -m gpu"""
import os

import numpy as np
import pytest
import torch

import ab_opt_b200
from oracle import weights, ipa_backward, epsnet_backward, transitions as T, geometry as G
from test_gpu_parity import build_model, cu, to64, DEV

pytestmark = pytest.mark.gpu


def close(name, got, ref32, ref64, rel=2e-5, factor=3.0):
    """err(cuda, fp64) <= factor * err(oracle fp32, fp64) + rel * max|ref|: the fp64 evaluation of the oracle arbitrates."""
    got, ref32 = got.detach().double().cpu(), ref32.double()
    scale = ref64.abs().max().item()
    e_got, e_ref = (got - ref64).abs().max().item(), (ref32 - ref64).abs().max().item()
    assert e_got <= factor * e_ref + rel * scale + 1e-12, f'{name}: cuda err {e_got:.3e}, oracle-fp32 err {e_ref:.3e}, scale {scale:.3e}'


@pytest.mark.parametrize('N,L,ragged', [(2, 24, True), (1, 72, False), (2, 130, True)])
def test_block_backward_vs_oracle(N, L, ragged):
    W = weights.make_state_dict(seed=5, num_layers=2, flavour='abdesign')
    model = build_model(W, 2, flavour='abdesign')
    inp = weights.synthetic_inputs(40 + L, N, L, gen_slices=((2, 6),), ragged=ragged)
    R, t = G.so3_exp(inp['v']), inp['p'] / 10.0
    g_out = torch.randn(N, L, 128, generator=torch.Generator().manual_seed(1))
    layer = 1
    pre = f'eps_net.encoder.blocks.{layer}.'
    gx32, gz32, gw32 = ipa_backward.ga_block_backward(W, pre, R, t, inp['res_feat'], inp['pair_feat'], inp['mask_res'], g_out)
    W64, i64 = weights.cast(W, torch.double), to64(inp)
    gx64, gz64, gw64 = ipa_backward.ga_block_backward(W64, pre, G.so3_exp(i64['v']), i64['p'] / 10.0, i64['res_feat'], i64['pair_feat'],
                                                      i64['mask_res'], g_out.double())
    ci = cu(inp)
    enc = model.eps_net.encoder
    gx, gz, gw = enc.block_backward(layer, R.to(DEV), t.to(DEV), ci['res_feat'], ci['pair_feat'], ci['mask_res'], g_out.to(DEV))
    close('d x', gx, gx32, gx64)
    close('d z', gz, gz32, gz64)
    assert (gz.cpu()[~inp['mask_res']] == 0).all()
    for k in gw32:
        close(k, gw[k[len('eps_net.encoder.'):]], gw32[k], gw64[k])


@pytest.mark.parametrize('flavour,obj', [('abdesign', 'pred_noise'), ('abdock', 'pred_x0'), ('abdock', 'pred_noise')])
def test_training_step_vs_oracle(flavour, obj):
    """loss_and_grads (one library call: forward + backward) against oracle.epsnet_backward.training_step in fp32 and fp64."""
    N, L, nl = 2, 40, 2
    W = weights.make_state_dict(seed=17, num_layers=nl, flavour=flavour)
    inp = weights.synthetic_inputs(31, N, L, gen_slices=((10, 22), (30, 34)), ragged=True)
    t = torch.tensor([88, 12])
    noise = T.draw_step_noise(N, L, torch.Generator().manual_seed(5))
    a = lambda WW, x, nz: epsnet_backward.training_step(WW, x['v'], x['p'], x['s'], x['res_feat'], x['pair_feat'], x['mask_generate'],
                                                        x['mask_res'], t, nz, flavour=flavour, obj=obj)
    l32, g32, r32, p32 = a(W, inp, noise)
    l64, g64, r64, p64 = a(weights.cast(W, torch.double), to64(inp), to64(noise))
    model = build_model(W, nl, flavour=flavour, obj=obj)
    ci = cu(inp)
    loss, grads, d_res, d_pair = model.loss_and_grads(ci['v'], ci['p'], ci['s'], ci['res_feat'], ci['pair_feat'], ci['mask_generate'],
                                                      ci['mask_res'], True, True, t=t.to(DEV), noise=cu(noise))
    assert sorted(loss) == sorted(l32)
    for k in l32:
        e_got, e_ref = abs(loss[k].double().item() - l64[k].item()), abs(l32[k].double().item() - l64[k].item())
        assert e_got <= 3 * e_ref + 2e-5 * max(1.0, abs(l64[k].item())), f'loss {k}: {loss[k].item()} vs {l64[k].item()}'
    assert sorted(grads) == sorted(g32)
    close('d res_feat', d_res, r32, r64)
    close('d pair_feat', d_pair, p32, p64)
    for k in g32:
        close(k, grads[k], g32[k], g64[k])


@pytest.mark.parametrize('obj', ['pred_x0', 'pred_noise'])
def test_training_step_against_reference_fixture(golden_dir, obj):
    """The gradients the UNMODIFIED REFERENCE computed with loss.backward() (AbDock flavour, tests/golden/train_backward.npz):
    grad-enabled losses, the gradient norm of all 71 parameters, ten full gradients, d / d res_feat, d / d pair_feat."""
    d = np.load(os.path.join(golden_dir, 'train_backward.npz'))
    g = {k: (torch.from_numpy(d[k]) if d[k].dtype.kind != 'U' and d[k].ndim else d[k]) for k in d.files}
    nl = int(g['num_layers'])
    W = weights.make_state_dict(seed=int(g['seed_w']), num_layers=nl, flavour='abdock')
    inp = weights.synthetic_inputs(int(g['seed_in']), int(g['N']), int(g['L']), gen_slices=((0, 5), (8, 10)), ragged=True)
    noise = T.draw_step_noise(int(g['N']), int(g['L']), torch.Generator().manual_seed(int(g['seed_noise'])))
    model = build_model(W, nl, flavour='abdock', obj=obj)
    ci = cu(inp)
    loss, grads, d_res, d_pair = model.loss_and_grads(ci['v'], ci['p'], ci['s'], ci['res_feat'], ci['pair_feat'], ci['mask_generate'],
                                                      ci['mask_res'], True, True, t=g['t'].to(DEV), noise=cu(noise))
    for k, v in loss.items():
        torch.testing.assert_close(v.cpu(), torch.as_tensor(g[f'{obj}_loss_{k}'].item()), rtol=1e-4, atol=1e-5, msg=lambda m, k=k: f'{k}: {m}')
    names = [str(x) for x in g[f'{obj}_param_names']]
    assert sorted(grads) == names
    norms = torch.stack([grads[k].double().norm().cpu() for k in names])
    torch.testing.assert_close(norms, g[f'{obj}_grad_norms'], rtol=2e-4, atol=1e-8)
    for key in d.files:
        if key.startswith(f'{obj}_grad_eps_net.'):
            want, got = g[key], grads[key[len(obj) + 6:]].cpu()
            assert (got - want).abs().max() <= 1e-4 * want.abs().max() + 1e-8, key
    for got, want in ((d_res.cpu(), g[f'{obj}_grad_res_feat']), (d_pair.cpu(), g[f'{obj}_grad_pair_feat'])):
        assert (got - want).abs().max() <= 1e-4 * want.abs().max()


def test_autograd_integration_and_loss_weights():
    """train.py's lines, unchanged: loss_dict = model(...); loss = sum_k w_k loss_k; loss.backward() -> .grad on every parameter and
    on res_feat / pair_feat, equal to the fused call with the same weights; validation (no_grad) carries no graph."""
    N, L, nl = 2, 32, 2
    W = weights.make_state_dict(seed=3, num_layers=nl, flavour='abdock')
    model = build_model(W, nl, flavour='abdock', obj='pred_x0').train()
    inp = weights.synthetic_inputs(7, N, L, gen_slices=((8, 16),), ragged=True)
    ci = cu(inp)
    noise = cu(T.draw_step_noise(N, L, torch.Generator().manual_seed(2)))
    t = torch.tensor([60, 9], device=DEV)
    wts = {'rot': 1.0, 'pos': 0.5, 'seq': 2.0, 'prmsd': 1.0, 'dist': 0.25}
    rf, pf = ci['res_feat'].clone().requires_grad_(True), ci['pair_feat'].clone().requires_grad_(True)
    loss = model(ci['v'], ci['p'], ci['s'], rf, pf, ci['mask_generate'], ci['mask_res'], True, True, t=t, noise=noise)
    assert sorted(loss) == ['dist', 'pos', 'prmsd', 'rot', 'seq'] and all(v.requires_grad for v in loss.values())
    total = sum(wts[k] * v for k, v in loss.items())
    total.backward()
    ref_loss, grads, d_res, d_pair = model.loss_and_grads(ci['v'], ci['p'], ci['s'], ci['res_feat'], ci['pair_feat'], ci['mask_generate'],
                                                          ci['mask_res'], True, True, t=t, noise=noise, loss_weights=wts)
    for k in ref_loss:
        torch.testing.assert_close(loss[k].detach(), ref_loss[k], rtol=1e-5, atol=1e-6)
    assert torch.equal(rf.grad, d_res) and torch.equal(pf.grad, d_pair)
    for k, p in model.named_parameters():
        assert p.grad is not None and torch.equal(p.grad, grads[k]), k
    assert any(p.grad.abs().max() > 0 for p in model.parameters())
    with torch.no_grad():
        val = model(ci['v'], ci['p'], ci['s'], ci['res_feat'], ci['pair_feat'], ci['mask_generate'], ci['mask_res'], True, True, t=t, noise=noise)
    assert not any(v.requires_grad for v in val.values())
