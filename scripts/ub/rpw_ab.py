"""A/B of the rows-per-warp setting of mixer_kernel / heads_kernel in their row-list launches (k_linear.cu): whole C2 samples
timed with CUDA events and the per-kernel breakdown, every setting in ONE process (the library reads ABOPT_RPW_* per call).
    python scripts/ub/rpw_ab.py [--config c2] [--samples 3] > gpurun_out/rpw_ab.jsonl"""
import argparse, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from ab_opt_b200 import _capi

ap = argparse.ArgumentParser()
ap.add_argument('--config', default='c2'); ap.add_argument('--samples', type=int, default=3)
ap.add_argument('--settings', default='8:8,2:4,4:4,2:2,4:2,8:8')
args = ap.parse_args()
cfg = dict(bench.CONFIGS[args.config])
dev = torch.device('cuda', 0)
model = bench.build_model(cfg, dev)
inp = bench.synthetic_batch(cfg, 1000, dev)
a = (inp['v'], inp['p'], inp['s'], inp['res_feat'], inp['pair_feat'], inp['mask_generate'], inp['mask_res'])
kw = dict(sample_structure=cfg['sample_structure'], sample_sequence=cfg['sample_sequence'])


def sample():
    torch.manual_seed(0)
    return model.sample(*a, **kw)


for _ in range(3):
    ref = sample()
torch.cuda.synchronize()
ref0 = [x.clone() for x in ref[0][:3]]
for st in args.settings.split(','):
    m, h = st.split(':')
    os.environ['ABOPT_RPW_MIXER'], os.environ['ABOPT_RPW_HEADS'] = m, h
    tr = sample()                                  # warm-up of this setting; same seed -> must reproduce the default's bits
    same = all(torch.equal(x, y) for x, y in zip(tr[0][:3], ref0))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.samples):
        tr = sample()
    e1.record(); torch.cuda.synchronize()
    _capi.profile_enable(True); sample(); torch.cuda.synchronize(); prof = _capi.profile_collect(); _capi.profile_enable(False)
    print(json.dumps({'mixer_rpw': int(m), 'heads_rpw': int(h), 'ms_per_sample': round(e0.elapsed_time(e1) / args.samples, 2),
                      'bit_equal_to_default': bool(same), 'mixer_ms': round(prof['mixer'][0], 3), 'heads_ms': round(prof['heads'][0], 3),
                      'config': args.config}), flush=True)
