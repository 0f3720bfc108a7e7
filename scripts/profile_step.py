"""A few reverse-diffusion steps of the bench workload, for ncu (launch list / --set full captures).

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python scripts/profile_step.py --config c2 --steps 2
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', default='c2')
    ap.add_argument('--steps', type=int, default=2)
    ap.add_argument('--B', type=int, default=None)
    args = ap.parse_args()
    cfg = dict(bench.CONFIGS[args.config])
    if args.B:
        cfg['B'] = args.B
    dev = torch.device('cuda', 0)
    model = bench.build_model(cfg, dev)
    inp = bench.synthetic_batch(cfg, 1000, dev)
    v, p, s = inp['v'], inp['p'], inp['s']
    for k in range(args.steps):
        out = model.reverse_step(100 - k, v, p, s, inp['res_feat'], inp['pair_feat'], inp['mask_generate'], inp['mask_res'],
                                 seed=7, sample_structure=cfg['sample_structure'], sample_sequence=cfg['sample_sequence'])
        v, p, s = out[0], out[1], out[2]
    torch.cuda.synchronize()
    print('ok', float(p.abs().mean()))
    import ctypes
    from ab_opt_b200 import _capi
    clk = (ctypes.c_longlong * 16)()
    _capi.check(_capi.lib().abopt_debug_clocks(clk))
    t0 = clk[0]
    names = ['start', 'a_full', 'b_full0', 'b_full1', 'tables', 'tmem_full', 'pass1', 'pass2', 'pass3', 'end']
    print('attn_logits CTA(0,0,0) timeline [cycles]:', {n: int(clk[k] - t0) for k, n in enumerate(names)})
    tn = ['start', 'out_transform done', 'LN1 + act stored', 'MLP done', 'end']
    print('outT_tail CTA 0 timeline [cycles]:', {n: int(clk[10 + k] - clk[10]) for k, n in enumerate(tn)})


if __name__ == '__main__':
    main()
