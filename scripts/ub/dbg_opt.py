"""Error statistics of the optimize replay test and of block taps at a small ragged shape, for the library selected by ABOPT_LIB."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import test_gpu_fullsize as T
from test_gpu_fullsize import *
W = weights.make_state_dict(seed=11, num_layers=2, flavour='abdock')
model = build_model(W, 2)
inp = weights.synthetic_inputs(21, 2, 40, gen_slices=((8, 18),), ragged=True)
N, L, T0 = 2, 40, 3
M = N * L
ci = cu(inp)
torch.manual_seed(321)
traj = model.optimize(ci['v'], ci['p'], ci['s'], T0, ci['res_feat'], ci['pair_feat'], ci['mask_generate'], ci['mask_res'], rng='torch')
torch.manual_seed(321)
def draws():
    return {'u': torch.randn(N, L, 3, device=DEV).cpu(), 'expo_ang': torch.empty(M, 8191, device=DEV).exponential_(1).cpu(),
            'unif_ang': torch.rand(M, device=DEV).cpu(), 'gauss_ang': torch.randn(M, device=DEV).cpu(),
            'z_pos': torch.randn(N, L, 3, device=DEV).cpu(), 'expo_seq': torch.empty(M, 20, device=DEV).exponential_(1).cpu()}
tape = {'init': draws()}
for t in range(T0, 0, -1):
    tape[t] = draws()
ref = sampler.sample(W, inp['v'], inp['p'], inp['s'], inp['res_feat'], inp['pair_feat'], inp['mask_generate'], inp['mask_res'],
                     obj='pred_x0', tape=tape, materialize=False, start_step=T0)
live = inp['mask_res']
for t in (3, 2, 1, 0):
    gap = (np.pi - ref[t][0].norm(dim=-1)).clamp_min(1e-9)
    err = (G.so3_exp(traj[t][0].cpu()) - G.so3_exp(ref[t][0])).abs().amax(dim=(-1, -2))
    bad = live & (gap > 0.1) & (err > 2e-4 + 1e-5 / gap ** 2)
    perr = (traj[t][1].cpu() - ref[t][1]).abs()[live].max()
    print(os.environ.get('ABOPT_LIB', 'new')[-30:], 't', t, 'seq equal', bool(torch.equal(traj[t][2].cpu()[live], ref[t][2][live])), 'pos err', float(perr),
          'rot err max', float(err[live].max()), 'bad', bad.nonzero().tolist(), [(float(err[i, j]), float(gap[i, j])) for i, j in bad.nonzero().tolist()])
