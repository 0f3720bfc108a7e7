"""Dev scratch: which kernel is not reproducible run-to-run at the full C2 batch?  (GPU box)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import test_gpu_fullsize as T
from test_gpu_parity import build_model
from oracle import weights, geometry as G

def nd(a, b):
    d = (a - b).abs()
    return f'{int((d > 0).sum())} entries differ, max {d.max().item():.3e}'

cfg = T.CONFIGS['c2']
B = int(os.environ.get('DBG_B', cfg['B']))
cfg = dict(cfg, B=B)
W = weights.make_state_dict(seed=29, num_layers=6, flavour='abdesign')
model = build_model(W, 6, flavour='abdesign', obj='pred_noise')
d = T.device_batch(cfg, 500)
R, t = G.so3_exp(d['v'].cpu()).to('cuda:0'), d['p'] / 10.0
enc = model.eps_net.encoder
x = d['res_feat']
for l in range(6):
    a1, f1 = enc.block_taps(l, R, t, x, d['pair_feat'], d['mask_res'])
    a2, f2 = enc.block_taps(l, R, t, x, d['pair_feat'], d['mask_res'])
    y1 = enc.blocks[l](R, t, x, d['pair_feat'], d['mask_res'])
    y2 = enc.blocks[l](R, t, x, d['pair_feat'], d['mask_res'])
    print(f'layer {l}: alpha {nd(a1, a2)} | feat pair {nd(f1[..., :768], f2[..., :768])} | feat node+pts {nd(f1[..., 768:], f2[..., 768:])} | out {nd(y1, y2)}', flush=True)
    del a1, a2, f1, f2
    x = y1
for rep in range(3):
    e1 = enc(R, t, d['res_feat'], d['pair_feat'], d['mask_res'])
    e2 = enc(R, t, d['res_feat'], d['pair_feat'], d['mask_res'])
    print('encoder (6 layers) run-to-run:', nd(e1, e2), '| vs chained single blocks:', nd(e1, x), flush=True)
