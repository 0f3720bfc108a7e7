"""Dev scratch: which workspace tensor of one GABlock call is not reproducible at the full C2 batch?  (GPU box)"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import test_gpu_fullsize as T
from test_gpu_parity import build_model
from ab_opt_b200 import _capi as C
from oracle import weights, geometry as G

cfg = dict(T.CONFIGS['c2'], B=int(os.environ.get('DBG_B', 64)))
W = weights.make_state_dict(seed=29, num_layers=6, flavour='abdesign')
model = build_model(W, 6, flavour='abdesign', obj='pred_noise')
d = T.device_batch(cfg, 500)
R, t = G.so3_exp(d['v'].cpu()).to('cuda:0'), d['p'] / 10.0
enc = model.eps_net.encoder
nm = enc.native()
names = ['QA', 'KB', 'rq', 'rk', 'VT', 'bias', 'alpha', 'feat']

def snap():
    out = {}
    for i, n in enumerate(names):
        buf = torch.empty(300_000_000 if n in ('bias', 'alpha', 'QA', 'KB', 'VT') else 40_000_000, device='cuda:0')
        k = ctypes.c_size_t()
        C.check(C.lib().abopt_debug_copy(nm.handle, i, C.ptr(buf), buf.numel(), ctypes.byref(k), C.stream_ptr(torch.device('cuda:0'))))
        torch.cuda.synchronize()
        out[n] = buf[:k.value].clone()
        del buf
    return out

for rep in range(3):
    enc.block_taps(0, R, t, d['res_feat'], d['pair_feat'], d['mask_res']); a = snap()
    enc.block_taps(0, R, t, d['res_feat'], d['pair_feat'], d['mask_res']); b = snap()
    line = []
    for n in names:
        df = (a[n] - b[n]).abs()
        nz = (df > 0).nonzero().flatten()
        line.append(f'{n}: {nz.numel()} differ' + (f' (first idx {int(nz[0])}, last {int(nz[-1])}, max {df.max().item():.2e})' if nz.numel() else ''))
    print(' | '.join(line), flush=True)
    if rep == 0:
        df = (a['alpha'] - b['alpha']).abs().reshape(cfg['B'], 12, 256, 256)
        bad = (df.amax(-1) > 0).nonzero()
        print('alpha rows differing (b, h, i):', bad[:12].tolist(), '... total rows', bad.shape[0])
        tiles = sorted({(int(x[0]), int(x[1]), int(x[2]) // 128) for x in bad})
        print('tiles (b, h, i0/128):', tiles[:40], len(tiles))
        # tile index -> CTA / order in its walk
        for (bb, hh, it) in tiles[:20]:
            tile = (bb * 12 + hh) * 2 + it
            print('  tile', tile, 'cta', tile % 148, 'n', tile // 148, 'rows', sorted({int(x[2]) for x in bad if (int(x[0]), int(x[1]), int(x[2]) // 128) == (bb, hh, it)})[:3], 'cols', (df[bb, hh] > 0).any(0).nonzero().flatten()[[0, -1]].tolist())
