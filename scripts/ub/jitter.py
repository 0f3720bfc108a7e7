"""Per-sample wall / device times of consecutive FullDPM.sample calls, with and without the nvidia-smi clock sampler."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
cfg = dict(bench.CONFIGS['c2'])
dev = torch.device('cuda', 0)
model = bench.build_model(cfg, dev)
inp = bench.synthetic_batch(cfg, 1000, dev)
a = (inp['v'], inp['p'], inp['s'], inp['res_feat'], inp['pair_feat'], inp['mask_generate'], inp['mask_res'])
def run(n, tag):
    out = []
    for k in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        traj = model.sample(*a, sample_structure=True, sample_sequence=True)
        e1.record(); torch.cuda.synchronize()
        out.append((round((time.perf_counter() - t0) * 1e3, 1), round(e0.elapsed_time(e1), 1)))
    print(tag, out, flush=True)
run(4, 'init')
run(8, 'plain')
cs = bench.ClockSampler(0)
run(8, 'with nvidia-smi sampler')
print(cs.stop())
run(8, 'plain again')
