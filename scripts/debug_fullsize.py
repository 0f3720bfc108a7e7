"""Dev scratch: where do full-size runs differ run-to-run / alone-vs-batch?  (run on the GPU box)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import test_gpu_fullsize as T
from test_gpu_parity import build_model
from oracle import weights

def diff(name, a, b):
    a, b = a.cpu(), b.cpu()
    if a.is_floating_point():
        d = (a - b).abs()
        bad = d.reshape(d.shape[0], d.shape[1], -1).amax(-1) > 0 if d.dim() > 2 else d > 0
        print(f'  {name}: differing entries {int((d > 0).sum())}, max {d.max().item():.3e}; rows {bad.nonzero()[:6].tolist()}')
    else:
        print(f'  {name}: differing entries {int((a != b).sum())}')

for name, nofocus in (('c2', True), ('c3', False)):
    cfg = T.CONFIGS[name]
    os.environ['ABOPT_NO_FOCUS'] = '1' if nofocus else '0'
    W = weights.make_state_dict(seed=29, num_layers=6, flavour=cfg['flavour'])
    model = build_model(W, 6, flavour=cfg['flavour'], obj=cfg['obj'])
    d = T.device_batch(cfg, 500)
    B, L = cfg['B'], cfg['L']
    nz = T.device_step_noise(B, L, 77)
    kw = dict(sample_structure=cfg['structure'], sample_sequence=cfg['sequence'])
    run = lambda x, n: model.reverse_step(61, x['v'], x['p'], x['s'], x['res_feat'], x['pair_feat'], x['mask_generate'], x['mask_res'], noise=n, **kw)
    g1 = run(d, nz); g2 = run(d, nz)
    print(name, 'nofocus', nofocus, 'run-to-run:')
    for i in range(len(g1)): diff(f'out{i}', g1[i], g2[i])
    al = run(T.pick(d, 0), T.noise_rows(nz, 0, L))
    print(' alone vs batch, complex 0:')
    for i in range(len(g1)): diff(f'out{i}', al[i], g1[i][0:1])
    # network outputs
    beta = W['trans_pos.var_sched.betas'][torch.full((B,), 61)].contiguous().to('cuda:0')
    a = lambda x, bt: (x['v'], x['p'] / 10.0, x['s'], x['res_feat'], x['pair_feat'], bt, x['mask_generate'], x['mask_res'])
    e1 = model.eps_net(*a(d, beta)); e2 = model.eps_net(*a(d, beta))
    print(' eps_net run-to-run:')
    for i in range(len(e1)): diff(f'net{i}', e1[i], e2[i])
    ea = model.eps_net(*a(T.pick(d, 0), beta[:1]))
    print(' eps_net alone vs batch, complex 0:')
    for i in range(len(e1)): diff(f'net{i}', ea[i], e1[i][0:1])
    gm = d['mask_generate'][0].nonzero().flatten().tolist()
    print(' generated rows of complex 0:', gm[:3], '...', 'v_t norms there', d['v'][0][gm[:4]].norm(dim=-1).tolist())

os.environ['ABOPT_NO_FOCUS'] = '0'
cfg = T.CONFIGS['c2']
W = weights.make_state_dict(seed=29, num_layers=6, flavour='abdesign')
model = build_model(W, 6, flavour='abdesign', obj='pred_noise')
d = T.device_batch(cfg, 600)
a = lambda x: (x['v'], x['p'], x['s'], 4, x['res_feat'], x['pair_feat'], x['mask_generate'], x['mask_res'])
w1 = model.optimize(*a(d), seed=1234); w2 = model.optimize(*a(d), seed=1234)
for t in (4, 3, 2, 1, 0):
    print('optimize run-to-run t =', t)
    for i in range(3): diff(f'f{i}', w1[t][i], w2[t][i])
