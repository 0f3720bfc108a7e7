"""Per-kernel launch times (CUDA events around every launch) over a few reverse steps of a bench configuration.
    [ABOPT_LIB=<variant .so>] python scripts/kbench.py [--config c2] [--steps 4] [--tag name]"""
import argparse, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from ab_opt_b200 import _capi

ap = argparse.ArgumentParser()
ap.add_argument('--config', default='c2'); ap.add_argument('--steps', type=int, default=4); ap.add_argument('--tag', default='')
ap.add_argument('--B', type=int, default=None)
args = ap.parse_args()
cfg = dict(bench.CONFIGS[args.config])
if args.B: cfg['B'] = args.B
dev = torch.device('cuda', 0)
model = bench.build_model(cfg, dev)
inp = bench.synthetic_batch(cfg, 1000, dev)
def run(k):
    v, p, s = inp['v'], inp['p'], inp['s']
    for i in range(k):
        o = model.reverse_step(100 - i, v, p, s, inp['res_feat'], inp['pair_feat'], inp['mask_generate'], inp['mask_res'], seed=7,
                               sample_structure=cfg['sample_structure'], sample_sequence=cfg['sample_sequence'])
        v, p, s = o[0], o[1], o[2]
    torch.cuda.synchronize()
run(2)
_capi.profile_enable(True); run(args.steps); prof = _capi.profile_collect(); _capi.profile_enable(False)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(args.steps); e1.record(); torch.cuda.synchronize()
print(args.tag or os.environ.get('ABOPT_LIB', 'default').split('/')[-2], f'step {e0.elapsed_time(e1) / args.steps:.3f} ms |',
      ' '.join(f'{k} {1e3 * v[0] / max(v[1], 1):.1f}us' for k, v in prof.items() if v[1]), flush=True)
