"""Phase timestamps of outT_tail_kernel (CTA 0, clock64): start, phase 1 (out_transform) done, LayerNorm 1 done, MLP done, end."""
import os, sys, ctypes, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from ab_opt_b200 import _capi
cfg = dict(bench.CONFIGS['c2']); dev = torch.device('cuda', 0)
model = bench.build_model(cfg, dev); inp = bench.synthetic_batch(cfg, 1000, dev)
os.environ['ABOPT_NO_FOCUS'] = '1'
for k in range(2):
    out = model.reverse_step(100 - k, inp['v'], inp['p'], inp['s'], inp['res_feat'], inp['pair_feat'], inp['mask_generate'], inp['mask_res'], seed=7)
torch.cuda.synchronize()
clk = (ctypes.c_longlong * 16)()
_capi.check(_capi.lib().abopt_debug_clocks(clk))
t0 = clk[10]
print({n: int(clk[10 + i] - t0) for i, n in enumerate(['start', 'phase 1 done', 'LN1 done', 'MLP done', 'end'])})
