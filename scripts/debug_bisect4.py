"""Dev scratch: per-chunk log-ratio of bad alpha rows (GPU box)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import test_gpu_fullsize as T
from test_gpu_parity import build_model
from oracle import weights, geometry as G
cfg = dict(T.CONFIGS['c2'])
W = weights.make_state_dict(seed=29, num_layers=6, flavour='abdesign')
model = build_model(W, 6, flavour='abdesign', obj='pred_noise')
d = T.device_batch(cfg, 500)
R, t = G.so3_exp(d['v'].cpu()).to('cuda:0'), d['p'] / 10.0
enc = model.eps_net.encoder
runs = [enc.block_taps(0, R, t, d['res_feat'], d['pair_feat'], d['mask_res'])[0] for _ in range(5)]
ref = torch.stack(runs).median(0).values
torch.set_printoptions(precision=3, linewidth=250, sci_mode=False)
for r, a in enumerate(runs[:3]):
    bad = ((a - ref).abs().amax(2) > 0).nonzero()
    print('run', r, 'bad rows', bad.shape[0])
    for (b, i, h) in bad[:5].tolist():
        x, y = a[b, i, :, h].double(), ref[b, i, :, h].double()
        lr = (x.clamp_min(1e-30) / y.clamp_min(1e-30)).log()
        ok = y > 1e-7
        per = []
        for c in range(8):
            sel = ok[c * 32:(c + 1) * 32]
            v = lr[c * 32:(c + 1) * 32][sel]
            per.append((round(v.mean().item(), 3), round((v.max() - v.min()).item(), 3)) if v.numel() else None)
        print(f'  b={b} h={h} i={i} (te {i % 128}): per 32-key chunk (mean log ratio, spread): {per}')
        c = 0
        print('     chunk 0 log-ratio by column:', lr[:32].float().cpu())
