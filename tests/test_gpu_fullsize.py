"""GPU parity at the BENCHMARKED shapes (BASELINE.json configs 2-5): the persistent kernels walk ~10 tiles per CTA there, which
the small-shape tests of test_gpu_parity.py never do.  Three kinds of evidence, all through the C ABI:

  (i)   independence -- complex k computed alone gives the SAME BITS as complex k inside the full batch (k = first, middle,
        last): tile walks, ring phase wraps, TMEM double buffering and per-warp row strides cannot leak between tiles;
  (ii)  the same complexes against the CPU oracle, arbitrated by its fp64 evaluation (err_cuda <= 2 err_oracle_fp32 + floor);
  (iii) trained weights (a slice of dock_single_cdr/250000.pt, fixture written by the unmodified reference) and a
        replayed-noise optimize() run against oracle.sampler.

Run on the B200 box:  pytest tests -m gpu
"""
import os

import numpy as np
import pytest
import torch

import ab_opt_b200
from oracle import weights, ipa, epsnet, sampler, transitions as T, geometry as G
from test_gpu_parity import assert_vs_fp64, build_model, cu, to64, DEV
from test_oracle_golden import trained_slice

pytestmark = pytest.mark.gpu

SIX = ((20, 30), (50, 60), (95, 105), (160, 170), (190, 200), (230, 240))
CONFIGS = {      # bench.py CONFIGS / SURVEY.md 8d
    'c2': dict(B=64, L=256, gen=((120, 136),), flavour='abdesign', obj='pred_noise', structure=True, sequence=True),
    'c3': dict(B=64, L=256, gen=((120, 136),), flavour='abdock', obj='pred_x0', structure=True, sequence=False),
    'c4': dict(B=32, L=320, gen=SIX, flavour='abdesign', obj='pred_noise', structure=True, sequence=True),
}
LAYERS = 6


def device_batch(cfg, seed, ragged_tail=True):
    """Seeded synthetic batch on the device.  The last complex is ragged (padding tail) so that the key mask
    and the masked-row paths run at full size as well."""
    B, L = cfg['B'], cfg['L']
    g = torch.Generator(device=DEV).manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g, device=DEV)
    v = G.uniform_so3_from_gauss4(rn(B, L, 4).cpu()).to(DEV)
    mask_generate = torch.zeros(B, L, dtype=torch.bool, device=DEV)
    for a, b in cfg['gen']:
        mask_generate[:, a:b] = True
    mask_res = torch.ones(B, L, dtype=torch.bool, device=DEV)
    s = torch.randint(0, 20, (B, L), generator=g, device=DEV)
    if ragged_tail:
        mask_res[-1, L - 9:] = False
        s[-1, L - 9:] = 21
        mask_generate &= mask_res
    return dict(v=v.contiguous(), p=rn(B, L, 3) * 10.0, s=s, res_feat=rn(B, L, 128), pair_feat=rn(B, L, L, 64),
                mask_generate=mask_generate, mask_res=mask_res)


def pick(d, k):
    return {n: t[k:k + 1].contiguous() for n, t in d.items()}


def noise_rows(nz, k, L):
    out = {}
    for n, t in nz.items():
        out[n] = (t[k:k + 1] if t.dim() == 3 else t[k * L:(k + 1) * L]).contiguous()
    return out


def device_step_noise(N, L, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    M = N * L
    return dict(u=torch.randn(N, L, 3, generator=g, device=DEV), expo_ang=torch.empty(M, 8191, device=DEV).exponential_(1, generator=g),
                unif_ang=torch.rand(M, generator=g, device=DEV), gauss_ang=torch.randn(M, generator=g, device=DEV),
                z_pos=torch.randn(N, L, 3, generator=g, device=DEV), expo_seq=torch.empty(M, 20, device=DEV).exponential_(1, generator=g))


def cpu(d):
    return {k: v.cpu() for k, v in d.items()}


@pytest.fixture(scope='module')
def models():
    cache = {}

    def get(flavour, obj):
        if (flavour, obj) not in cache:
            W = weights.make_state_dict(seed=29, num_layers=LAYERS, flavour=flavour)
            cache[(flavour, obj)] = (W, build_model(W, LAYERS, flavour=flavour, obj=obj))
        return cache[(flavour, obj)]
    return get


@pytest.fixture(autouse=True)
def _restore_focus_env():
    old = os.environ.get('ABOPT_NO_FOCUS')
    yield
    if old is None:
        os.environ.pop('ABOPT_NO_FOCUS', None)
    else:
        os.environ['ABOPT_NO_FOCUS'] = old


def ks_of(B):
    return (0, B // 2 - 1, B - 1)


# ------------------------------------------------------------------------------------------ (i) + (ii): one GABlock
@pytest.mark.parametrize('name,layer', [('c2', 0), ('c2', 5), ('c4', 3)])
def test_block_taps_full_batch(models, name, layer):
    """alpha (N,L,L,12) and the 1824-wide aggregate of one GABlock over the FULL batch: ~10 tiles per CTA in the logits /
    aggregation kernels, ~9 query rows per warp in pair_stream_kernel."""
    cfg = CONFIGS[name]
    W, model = models(cfg['flavour'], cfg['obj'])
    d = device_batch(cfg, 300 + layer)
    B, L = cfg['B'], cfg['L']
    R, t = G.so3_exp(d['v'].cpu()).to(DEV), d['p'] / 10.0
    enc = model.eps_net.encoder
    alpha, feat = enc.block_taps(layer, R, t, d['res_feat'], d['pair_feat'], d['mask_res'])
    out = enc.blocks[layer](R, t, d['res_feat'], d['pair_feat'], d['mask_res'])
    assert torch.isfinite(alpha).all() and torch.isfinite(feat).all() and torch.isfinite(out).all()
    pre = f'eps_net.encoder.blocks.{layer}.'
    W64 = weights.cast(W, torch.double)
    for k in ks_of(B):
        a1, f1 = enc.block_taps(layer, R[k:k + 1].contiguous(), t[k:k + 1].contiguous(), d['res_feat'][k:k + 1].contiguous(),
                                d['pair_feat'][k:k + 1].contiguous(), d['mask_res'][k:k + 1].contiguous())
        o1 = enc.blocks[layer](R[k:k + 1].contiguous(), t[k:k + 1].contiguous(), d['res_feat'][k:k + 1].contiguous(),
                               d['pair_feat'][k:k + 1].contiguous(), d['mask_res'][k:k + 1].contiguous())
        assert torch.equal(a1[0], alpha[k]), f'alpha of complex {k} depends on the batch it is computed in'
        mr = d['mask_res'][k]
        assert torch.equal(f1[0][mr], feat[k][mr]), f'aggregate of complex {k} depends on the batch'
        assert torch.equal(o1[0], out[k]), f'block output of complex {k} depends on the batch'
        c = cpu(pick(d, k))
        Rk, tk = G.so3_exp(c['v']), c['p'] / 10.0
        o32, p32 = ipa.ga_block(W, pre, Rk, tk, c['res_feat'], c['pair_feat'], c['mask_res'], materialize=False, return_parts=True)
        c64 = to64(c)
        o64, p64 = ipa.ga_block(W64, pre, G.so3_exp(c64['v']), c64['p'] / 10.0, c64['res_feat'], c64['pair_feat'], c64['mask_res'],
                                materialize=False, return_parts=True)
        assert_vs_fp64(f'alpha[{k}]', alpha[k:k + 1], p32['alpha'], p64['alpha'], 5e-6)
        m = c['mask_res']
        fg, f32, f64 = feat[k:k + 1].cpu()[m], p32['feat'][m], p64['feat'][m]
        assert_vs_fp64(f'feat[{k}]', fg[:, :1536], f32[:, :1536], f64[:, :1536], 3e-5)
        assert_vs_fp64(f'feat.dir[{k}]', fg[:, 1536:], f32[:, 1536:], f64[:, 1536:], 2e-4)
        assert_vs_fp64(f'x_out[{k}]', out[k:k + 1], o32, o64, 3e-5)
        a = alpha[k].cpu()
        torch.testing.assert_close(a[m[0]].sum(1), torch.ones_like(a[m[0]].sum(1)), rtol=0, atol=1e-5)
        if (~m).any():
            assert a[~m[0]].abs().max() == 0 and a.transpose(0, 1)[~m[0]].abs().max() == 0


# ------------------------------------------------------------------------------------------ (i) + (ii): EpsilonNet, 6 layers
@pytest.mark.parametrize('name', ['c2', 'c3', 'c4'])
def test_eps_net_full_batch(models, name):
    cfg = CONFIGS[name]
    W, model = models(cfg['flavour'], cfg['obj'])
    d = device_batch(cfg, 400)
    B, L = cfg['B'], cfg['L']
    tsteps = torch.randint(1, 100, (B,), generator=torch.Generator().manual_seed(4))
    beta = W['trans_pos.var_sched.betas'][tsteps].contiguous()
    a = lambda x, bt: (x['v'], x['p'] / 10.0, x['s'], x['res_feat'], x['pair_feat'], bt, x['mask_generate'], x['mask_res'])
    got = model.eps_net(*a(d, beta.to(DEV)))
    W64 = weights.cast(W, torch.double)
    floors = (None, 5e-6, 5e-6, 5e-7, 5e-6)
    for k in ks_of(B):
        alone = model.eps_net(*a(pick(d, k), beta[k:k + 1].to(DEV)))
        for i in range(len(got)):
            assert torch.equal(alone[i][0], got[i][k]), f'output {i} of complex {k} depends on the batch'
        c = cpu(pick(d, k))
        o32 = epsnet.eps_net(W, *a(c, beta[k:k + 1]), materialize=False)
        o64 = epsnet.eps_net(W64, *a(to64(c), beta[k:k + 1].double()), materialize=False)
        for i in range(1, len(got)):
            assert_vs_fp64(f'{name} out{i}[{k}]', got[i][k:k + 1], o32[i], o64[i], floors[i])


# ------------------------------------------------------------------------------------------ one reverse step, replayed noise
@pytest.mark.parametrize('name,no_focus', [('c2', False), ('c2', True), ('c3', False), ('c4', False)])
def test_reverse_step_full_batch(models, name, no_focus):
    """One iteration of the sampling loop at the benchmarked shape (focus mode on and off for C2): independence bit for bit,
    then positions 1e-4 relative, rotations arbitrated by the fp64 oracle, sampled amino-acid indices BIT-EXACT."""
    cfg = CONFIGS[name]
    os.environ['ABOPT_NO_FOCUS'] = '1' if no_focus else '0'
    W, model = models(cfg['flavour'], cfg['obj'])
    d = device_batch(cfg, 500)
    B, L = cfg['B'], cfg['L']
    t = 61
    nz = device_step_noise(B, L, 77)
    kw = dict(sample_structure=cfg['structure'], sample_sequence=cfg['sequence'])
    run = lambda x, n: model.reverse_step(t, x['v'], x['p'], x['s'], x['res_feat'], x['pair_feat'], x['mask_generate'], x['mask_res'],
                                          noise=n, **kw)
    got = run(d, nz)
    assert all(torch.isfinite(x).all() for x in got if x.is_floating_point())
    keep = ~d['mask_generate']
    assert torch.equal(got[0][keep], d['v'][keep]) and torch.equal(got[1][keep], d['p'][keep])
    W64 = weights.cast(W, torch.double)
    for k in ks_of(B):
        alone = run(pick(d, k), noise_rows(nz, k, L))
        for i in range(len(got)):
            assert torch.equal(alone[i][0], got[i][k]), f'{name}: output {i} of complex {k} depends on the batch (focus off={no_focus})'
        c, cn = cpu(pick(d, k)), cpu(noise_rows(nz, k, L))
        ref = sampler.reverse_step(W, t, c['v'], c['p'] / 10.0, c['s'], c['res_feat'], c['pair_feat'], c['mask_generate'], c['mask_res'],
                                   cn, obj=cfg['obj'], materialize=False)
        c64 = to64(c)
        ref64 = sampler.reverse_step(W64, t, c64['v'], c64['p'] / 10.0, c64['s'], c64['res_feat'], c64['pair_feat'], c64['mask_generate'],
                                     c64['mask_res'], to64(cn), obj=cfg['obj'], materialize=False)
        v_o, p_o, s_o = got[0][k:k + 1].cpu(), got[1][k:k + 1].cpu(), got[2][k:k + 1].cpu()
        live = c['mask_res']
        torch.testing.assert_close(p_o[live], (ref['p_next'] * 10.0)[live], rtol=1e-4, atol=2e-4)
        R64 = G.so3_exp(ref64['v_next'])
        e_cuda = (G.so3_exp(v_o.double()) - R64).abs().amax(dim=(-1, -2))
        e_o32 = (G.so3_exp(ref['v_next'].double()) - R64).abs().amax(dim=(-1, -2))
        gap = (np.pi - ref64['v_next'].norm(dim=-1)).clamp_min(1e-9)
        ok = live & (gap > 0.05) & (e_o32 < 1e-3)
        bad = ok & (e_cuda > 4 * e_o32 + 3e-5 + 3e-6 / gap ** 2)
        assert ok[c['mask_generate']].float().mean() > 0.7
        assert not bad.any(), f'{name} v_next[{k}]: cuda {e_cuda[bad].max().item():.3e} oracle32 {e_o32[bad].max().item():.3e}'
        if cfg['sequence']:
            assert torch.equal(s_o[live], ref['s_next'][live]), f'{name}: sampled amino acids of complex {k} differ from the oracle'
        else:
            assert torch.equal(s_o, c['s'])
        if cfg['flavour'] == 'abdock':
            torch.testing.assert_close(got[3][k:k + 1].cpu(), ref['prmsd'], rtol=1e-4, atol=2e-4)


# ------------------------------------------------------------------------------------------ sampling loop properties at C2
def test_sample_loop_c2_sharded_equals_whole():
    """Philox mode at B=64, L=256, 6 layers, 4 steps of optimize(): deterministic for a seed, context untouched, and a slice of
    the batch run on its own with batch_offset reproduces the whole-batch run BIT FOR BIT (what the N-rank run relies on)."""
    cfg = CONFIGS['c2']
    W = weights.make_state_dict(seed=29, num_layers=LAYERS, flavour='abdesign')
    model = build_model(W, LAYERS, flavour='abdesign', obj='pred_noise')
    d = device_batch(cfg, 600)
    a = lambda x: (x['v'], x['p'], x['s'], 4, x['res_feat'], x['pair_feat'], x['mask_generate'], x['mask_res'])
    whole = model.optimize(*a(d), seed=1234)
    again = model.optimize(*a(d), seed=1234)
    other = model.optimize(*a(d), seed=1235)
    for i in range(3):
        assert torch.equal(whole[0][i], again[0][i]) and torch.equal(whole[2][i], again[2][i].to(whole[2][i].device))
    assert not torch.equal(whole[0][1], other[0][1])
    keep = ~d['mask_generate']
    assert torch.equal(whole[0][0][keep], d['v'][keep]) and torch.equal(whole[0][2][keep & d['mask_res']], d['s'][keep & d['mask_res']])
    lo, hi = 40, 56                                      # "rank 5 of 8" in a contiguous split of 64... any slice will do
    part = {n: t[lo:hi].contiguous() for n, t in d.items()}
    sl = model.optimize(*a(part), seed=1234, batch_offset=lo, batch_total=cfg['B'])
    for t_ in (4, 2, 0):
        for i in range(3):
            assert torch.equal(sl[t_][i].cpu(), whole[t_][i][lo:hi].cpu()), f'slice differs from the whole batch at t={t_}, field {i}'


def test_optimize_replayed_noise_matches_oracle():
    """FullDPM.optimize(opt_step=3) in parity mode (rng='torch'): the same CUDA draws replayed through oracle.sampler.sample
    (start_step=3) reproduce the noised start and the three reverse steps (dpm_full.py:304-367)."""
    W = weights.make_state_dict(seed=11, num_layers=2, flavour='abdock')
    model = build_model(W, 2)
    inp = weights.synthetic_inputs(21, 2, 40, gen_slices=((8, 18),), ragged=True)
    N, L, T0 = 2, 40, 3
    M = N * L
    ci = cu(inp)
    torch.manual_seed(321)
    traj = model.optimize(ci['v'], ci['p'], ci['s'], T0, ci['res_feat'], ci['pair_feat'], ci['mask_generate'], ci['mask_res'], rng='torch')
    torch.manual_seed(321)

    def draws():      # the reference's order: so3.py:143,123,126,131 ; transition.py:74/95 ; transition.py:199/179
        return {'u': torch.randn(N, L, 3, device=DEV).cpu(), 'expo_ang': torch.empty(M, 8191, device=DEV).exponential_(1).cpu(),
                'unif_ang': torch.rand(M, device=DEV).cpu(), 'gauss_ang': torch.randn(M, device=DEV).cpu(),
                'z_pos': torch.randn(N, L, 3, device=DEV).cpu(), 'expo_seq': torch.empty(M, 20, device=DEV).exponential_(1).cpu()}
    tape = {'init': draws()}
    for t in range(T0, 0, -1):
        tape[t] = draws()
    ref = sampler.sample(W, inp['v'], inp['p'], inp['s'], inp['res_feat'], inp['pair_feat'], inp['mask_generate'], inp['mask_res'],
                         obj='pred_x0', tape=tape, materialize=False, start_step=T0)
    assert sorted(traj) == [0, 1, 2, 3] and isinstance(traj[2], tuple)
    live = inp['mask_res']
    # A rotation that passes within 0.1 rad of pi is allowed to deviate there (the reference's log map is ill-conditioned near pi,
    # SURVEY finding 4: rounding noise of 1e-7 becomes 1e-5 / gap^2), and it STAYS deviated in the steps after it: such a residue
    # is not compared again (its neighbours still are, and so are all positions and sequences).
    tainted = torch.zeros_like(live)
    for t in (3, 2, 1, 0):
        assert torch.equal(traj[t][2].cpu()[live], ref[t][2][live]), f'sequence differs at t={t}'
        torch.testing.assert_close(traj[t][1].cpu()[live], ref[t][1][live], rtol=1e-4, atol=2e-3)
        gap = (np.pi - ref[t][0].norm(dim=-1)).clamp_min(1e-9)
        err = (G.so3_exp(traj[t][0].cpu()) - G.so3_exp(ref[t][0])).abs().amax(dim=(-1, -2))
        assert not (live & ~tainted & (gap > 0.1) & (err > 2e-4 + 1e-5 / gap ** 2)).any(), f'rotations differ at t={t}'
        tainted |= (gap <= 0.1) & (err > 2e-4)
    assert int(tainted.sum()) <= 2, 'too many residues excluded'
    torch.testing.assert_close(traj[2][3].cpu(), ref[2][3], rtol=1e-4, atol=1e-4)       # pRMSD
    torch.testing.assert_close(traj[2][4].cpu(), ref[2][4], rtol=1e-4, atol=1e-5)       # perplexity (no mask in optimize)


@pytest.mark.parametrize('name,B', [('c2', 16), ('c4', 5), ('c3', 6)])
def test_context_cache_matches_full_stream(name, B):
    """Inside the sampling loop the first GABlock reuses the context part of its pair aggregate (k_pair.cu: ctx_delta_kernel)
    instead of streaming all of z: the same numbers up to rounding (fp32 reassociation in ctx_delta_kernel against the 3xTF32
    accumulation of pair_stream_kernel).  Philox mode, same seed, with the cache and with ABOPT_NO_CTXCACHE=1; one reverse step
    apart the states agree to rounding, three steps apart to what rounding grows into (measured 5.1e-3 A at C2)."""
    cfg = dict(CONFIGS[name]); cfg['B'] = B
    W = weights.make_state_dict(seed=31, num_layers=3, flavour=cfg['flavour'])
    model = build_model(W, 3, flavour=cfg['flavour'], obj=cfg['obj'])
    d = device_batch(cfg, 700)
    d['mask_res'][0, 3:7] = False                         # masked context queries / keys in the first complex as well
    d['mask_generate'] &= d['mask_res']
    a = lambda x: (x['v'], x['p'], x['s'], 3, x['res_feat'], x['pair_feat'], x['mask_generate'], x['mask_res'])
    kw = dict(seed=77, sample_structure=cfg['structure'], sample_sequence=cfg['sequence'])
    old = os.environ.get('ABOPT_NO_CTXCACHE')
    try:
        os.environ.pop('ABOPT_NO_CTXCACHE', None)
        fast = model.optimize(*a(d), **kw)
        os.environ['ABOPT_NO_CTXCACHE'] = '1'
        full = model.optimize(*a(d), **kw)
    finally:
        if old is None:
            os.environ.pop('ABOPT_NO_CTXCACHE', None)
        else:
            os.environ['ABOPT_NO_CTXCACHE'] = old
    live = d['mask_res'].cpu()
    for t_, tol in ((3, 0.0), (2, 2e-4), (0, 1e-2)):
        pf, pu = fast[t_][1].cpu()[live], full[t_][1].cpu()[live]
        assert torch.isfinite(pf).all()
        assert (pf - pu).abs().max() <= tol, f'positions differ at t={t_}: {(pf - pu).abs().max()}'
        rf, ru = G.so3_exp(fast[t_][0].cpu())[live], G.so3_exp(full[t_][0].cpu())[live]
        bad = ((rf - ru).abs().amax(dim=(-1, -2)) > 10 * tol + 1e-6)
        assert int(bad.sum()) <= (0 if t_ == 3 else 2), f'rotations differ at t={t_}'
        flips = int((fast[t_][2].cpu()[live] != full[t_][2].cpu()[live]).sum())
        assert flips <= (0 if t_ == 3 else 2), f'{flips} sequence flips at t={t_}'


# ------------------------------------------------------------------------------------------ (iii) trained weights
def test_trained_checkpoint_slice(golden_dir):
    """Blocks 0-1, mixer and heads of the reference's dock_single_cdr/250000.pt on features produced by the checkpoint's own
    embeddings: the 3xTF32 tensor-core path at trained activation scales, against what the UNMODIFIED reference computed."""
    W, g = trained_slice(golden_dir)
    model = build_model(W, g['num_layers'])
    N, L = g['N'], g['L']
    d = cu({k: g[k] for k in ('v', 'p', 's', 'res_feat', 'pair_feat', 'mask_generate', 'mask_res')})
    R, t = G.so3_exp(g['v']).to(DEV), d['p'] / 10.0
    enc = model.eps_net.encoder
    alpha, feat = enc.block_taps(0, R, t, d['res_feat'], d['pair_feat'], d['mask_res'])
    torch.testing.assert_close(alpha.cpu(), g['alpha'], rtol=1e-4, atol=2e-6)
    mr = g['mask_res']
    torch.testing.assert_close(feat.cpu()[mr][:, :1536], g['feat'][mr][:, :1536], rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(enc.blocks[0](R, t, d['res_feat'], d['pair_feat'], d['mask_res']).cpu(), g['x_out'], rtol=1e-4, atol=5e-5)
    torch.testing.assert_close(enc(R, t, d['res_feat'], d['pair_feat'], d['mask_res']).cpu(), g['enc_out'], rtol=1e-4, atol=5e-5)
    beta = W['trans_pos.var_sched.betas'][g['t']].expand(N).contiguous()
    got = model.eps_net(d['v'], t, d['s'], d['res_feat'], d['pair_feat'], beta.to(DEV), d['mask_generate'], d['mask_res'])
    W64 = weights.cast(W, torch.double)
    g64 = to64({k: g[k] for k in ('v', 'p', 's', 'res_feat', 'pair_feat', 'mask_generate', 'mask_res')})
    o64 = epsnet.eps_net(W64, g64['v'], g64['p'] / 10.0, g64['s'], g64['res_feat'], g64['pair_feat'], beta.double(), g64['mask_generate'],
                         g64['mask_res'], materialize=False)
    # floors: a few ulp of the 3xTF32 products at trained activation scales (|x| up to 4, logits up to 20)
    for i, (key, floor) in enumerate((('v_next', None), ('R_next', 1e-5), ('eps_pos', 5e-6), ('c_denoised', 2e-5), ('prmsd_logits', 3e-5))):
        if floor is None:
            continue
        assert_vs_fp64(key, got[i], g[key], o64[i], floor)             # "ref32" here IS the unmodified reference's fp32 result


# ------------------------------------------------------------------------------------------ training forward at C5 size
@torch.no_grad()      # the validate() evaluation; the grad-enabled step is covered by test_gpu_backward.py
def test_training_forward_c5_is_the_mean_of_its_halves():
    """FullDPM.forward at B=128, L=256, 6 layers: every loss is a masked mean with the same number of generated residues per
    complex, so the loss of the batch equals the mean of the losses of its two halves (a size-independent property), and a
    two-complex slice agrees with the oracle."""
    from oracle import training
    cfg = dict(B=128, L=256, gen=((120, 136),))
    W = weights.make_state_dict(seed=29, num_layers=LAYERS, flavour='abdesign')
    model = build_model(W, LAYERS, flavour='abdesign', obj='pred_noise')
    d = device_batch(cfg, 700, ragged_tail=False)
    B, L = cfg['B'], cfg['L']
    t = torch.randint(1, 100, (B,), generator=torch.Generator().manual_seed(8)).to(DEV)
    nz = device_step_noise(B, L, 88)
    f = lambda x, tt, n: model(x['v'], x['p'], x['s'], x['res_feat'], x['pair_feat'], x['mask_generate'], x['mask_res'], True, True,
                               t=tt, noise=n)
    whole = f(d, t, nz)
    h = B // 2
    halves = []
    for lo in (0, h):
        part = {n: x[lo:lo + h].contiguous() for n, x in d.items()}
        pn = {n: (x[lo:lo + h] if x.dim() == 3 else x[lo * L:(lo + h) * L]).contiguous() for n, x in nz.items()}
        halves.append(f(part, t[lo:lo + h].contiguous(), pn))
    for k in whole:
        want = 0.5 * (halves[0][k].double() + halves[1][k].double())
        torch.testing.assert_close(whole[k].double(), want, rtol=2e-5, atol=1e-6, msg=lambda m, k=k: f'{k}: {m}')
    two = {n: x[B - 2:].contiguous() for n, x in d.items()}
    n2 = {n: (x[B - 2:] if x.dim() == 3 else x[(B - 2) * L:]).contiguous() for n, x in nz.items()}
    got = f(two, t[B - 2:].contiguous(), n2)
    c, cn = cpu(two), cpu(n2)
    la = lambda WW, x, n: training.loss_forward(WW, x['v'], x['p'], x['s'], x['res_feat'], x['pair_feat'], x['mask_generate'], x['mask_res'],
                                                True, True, t[B - 2:].cpu(), n, flavour='abdesign', obj='pred_noise')
    ref, ref64 = la(W, c, cn), la(weights.cast(W, torch.double), to64(c), to64(cn))
    # the fp64 oracle arbitrates: noised rotations near pi make the fp32 oracle itself uncertain at the 1e-3 level on `rot`
    for k in ref:
        e_got, e_ref = abs(got[k].double().item() - ref64[k].item()), abs(ref[k].double().item() - ref64[k].item())
        assert e_got <= 2 * e_ref + 2e-5 * max(1.0, abs(ref64[k].item())), f'{k}: cuda {got[k].item()} oracle32 {ref[k].item()} oracle64 {ref64[k].item()}'


# ------------------------------------------------------------------------------------------ encode + sample from atoms
def test_design_entry_points_from_atoms():
    """DiffusionAntibodyDesign.sample / optimize (models/diffab.py:115-171) as one device-resident call: the trajectory's
    context rows carry v_0 = log(construct_3d_basis(CA, C, N)), p_0 = CA, s_0 = aa of the oracle; the composed path (module
    mirrors + FullDPM with the same seed) agrees on the first frames; host and device entry points give the same bits."""
    from oracle import pair_embed as PE
    C = ab_opt_b200._capi
    N, L, A = 3, 40, 15
    cfg = dict(res_feat_dim=128, pair_feat_dim=64, num_bins=40, dist_min=0.5, dist_max=19.5,
               diffusion=dict(num_steps=100, eps_net_opt=dict(num_layers=2), obj='pred_x0'))
    model = ab_opt_b200.DiffusionAntibodyDesign(cfg)
    model.diffusion.load_state_dict(weights.make_state_dict(seed=11, num_layers=2, flavour='abdock'), strict=True)
    model.pair_embed.load_state_dict(PE.make_state_dict(2, A), strict=True)
    model.residue_embed.load_state_dict(PE.make_residue_state_dict(3, A), strict=True)
    model = model.to(DEV).eval()
    cx = PE.synthetic_complex(8, N, L)
    gen = ~cx['context_mask'] & cx['mask_atoms'][:, :, 1]
    batch = dict(aa=cx['aa'].clamp(max=19), res_nb=cx['res_nb'], chain_nb=cx['chain_nb'], pos_heavyatom=cx['pos_atoms'],
                 mask_heavyatom=cx['mask_atoms'], fragment_type=torch.randint(1, 4, (N, L), generator=torch.Generator().manual_seed(2)),
                 generate_flag=gen, mask=cx['mask_atoms'][:, :, 1].clone())
    dbatch = cu(batch)
    T0 = 4
    traj = model.optimize(dict(dbatch), T0, {'sample_structure': True, 'sample_sequence': True, 'seed': 99})
    again = model.optimize(dict(dbatch), T0, {'sample_structure': True, 'sample_sequence': True, 'seed': 99})
    assert sorted(traj) == list(range(T0 + 1))
    for i in range(3):
        assert torch.equal(traj[0][i], again[0][i])
    # context rows of every frame = the encoded input state
    R0 = PE.backbone_frames(cx['pos_atoms'][:, :, 1], cx['pos_atoms'][:, :, 2], cx['pos_atoms'][:, :, 0])
    v0 = G.so3_log(R0)
    keep = ~gen & batch['mask']
    gap = (np.pi - v0.norm(dim=-1)).clamp_min(1e-9)                   # the log map is ill-conditioned near pi (SURVEY finding 4)
    err = (G.so3_exp(traj[T0][0].cpu()) - G.so3_exp(v0)).abs().amax(dim=(-1, -2))
    assert not (keep & (gap > 0.05) & (err > 2e-5 + 3e-6 / gap ** 2)).any(), err[keep].max()
    # positions are re-normalised / un-normalised on every step (dpm_full.py:276,297): equal to rounding, not bit for bit
    torch.testing.assert_close(traj[0][1].cpu()[keep], cx['pos_atoms'][:, :, 1][keep], rtol=1e-5, atol=1e-4)
    assert torch.equal(traj[0][2].cpu()[keep], batch['aa'][keep])
    assert torch.isfinite(traj[0][1]).all() and (traj[0][0].cpu()[gen] != v0[gen]).any()
    # the composed path of the reference (encode, then diffusion.optimize) with the same seed: same noised start, close after one step
    res_feat, pair_feat, R, p = model.encode(dict(dbatch), True, True)
    comp = model.diffusion.optimize(model._so3vec(R), p, dbatch['aa'], T0, res_feat, pair_feat, dbatch['generate_flag'], dbatch['mask'],
                                    seed=99)
    live = batch['mask']
    assert torch.equal(comp[T0][2][live], traj[T0][2][live])
    torch.testing.assert_close(comp[T0][1][live], traj[T0][1][live], rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(comp[T0 - 1][1][live], traj[T0 - 1][1][live], rtol=1e-3, atol=5e-3)
    # host entry point: same bits as the device one
    nm, pe, re_ = model.diffusion.native(), model.pair_embed.native(), model.residue_embed.native()
    flags = C.SAMPLE_STRUCTURE | C.SAMPLE_SEQUENCE | C.KEEP_TRAJECTORY
    tv, tp, ts = torch.empty(T0 + 1, N, L, 3), torch.empty(T0 + 1, N, L, 3), torch.empty(T0 + 1, N, L, dtype=torch.int64)
    pr, pl = torch.empty(T0 + 1, N), torch.empty(T0 + 1, N)
    h = {k: v.contiguous() for k, v in batch.items()}
    C.check(C.lib().abopt_design_host(nm.handle, pe.handle, re_.handle, N, L, A, C.ptr(h['aa']), C.ptr(h['res_nb']), C.ptr(h['chain_nb']),
                                      C.ptr(h['pos_heavyatom']), C.ptr(h['mask_heavyatom']), C.ptr(h['fragment_type']),
                                      C.ptr(h['generate_flag']), C.ptr(h['mask']), flags, T0, 99, C.ptr(tv), C.ptr(tp), C.ptr(ts),
                                      C.ptr(pr), C.ptr(pl)))
    assert torch.equal(tv[0], traj[0][0].cpu()) and torch.equal(tp[0], traj[0][1].cpu()) and torch.equal(ts[0], traj[0][2].cpu())
    assert torch.equal(tv[2], traj[2][0]) and torch.equal(ts[T0], traj[T0][2])
    # sample() through the reference-shaped call
    out = model.sample(dict(dbatch), {'sample_structure': True, 'sample_sequence': True, 'contig': '', 'seed': 5})
    assert sorted(out) == list(range(101)) and isinstance(out[3], list) and len(out[3]) == 5 and out[0][0].is_cuda
