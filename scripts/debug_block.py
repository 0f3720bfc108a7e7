import sys, torch
sys.path.insert(0, '.')
import ab_opt_b200
from oracle import weights, ipa, geometry as G
DEV='cuda:0'
W = weights.make_state_dict(seed=5, num_layers=1, flavour='abdesign')
m = ab_opt_b200.FullDPMAbDesign(128, 64, 100, eps_net_opt=dict(num_layers=1)); m.load_state_dict(W); m = m.to(DEV)
for (N, L, rag) in [(2, 24, True), (2, 64, False), (3, 100, True)]:
    inp = weights.synthetic_inputs(100 + L, N, L, gen_slices=((2, 6),), ragged=rag)
    R, t = G.so3_exp(inp['v']), inp['p'] / 10.0
    o32, parts = ipa.ga_block(W, 'eps_net.encoder.blocks.0.', R, t, inp['res_feat'], inp['pair_feat'], inp['mask_res'], materialize=False, return_parts=True)
    ci = {k: v.to(DEV) for k, v in inp.items()}
    enc = m.eps_net.encoder
    alpha, feat = enc.block_taps(0, R.to(DEV), t.to(DEV), ci['res_feat'], ci['pair_feat'], ci['mask_res'])
    out = enc.blocks[0](R.to(DEV), t.to(DEV), ci['res_feat'], ci['pair_feat'], ci['mask_res'])
    mr = inp['mask_res']
    def err(a, b): return (a - b).abs().max().item()
    print(f'N={N} L={L}: alpha {err(alpha.cpu(), parts["alpha"]):.3e}')
    f, fr = feat.cpu()[mr], parts['feat'][mr]
    for nm, a, b in (('p2n', 0, 768), ('node', 768, 1152), ('pts', 1152, 1440), ('dist', 1440, 1536), ('dir', 1536, 1824)):
        print(f'   {nm}: {err(f[:, a:b], fr[:, a:b]):.3e} (scale {fr[:, a:b].abs().max().item():.2f})')
    print(f'   out: {err(out.cpu(), o32):.3e}')
    # per-head alpha error
    ea = (alpha.cpu() - parts['alpha']).abs().amax(dim=(0, 1, 2))
    print('   alpha err per head', [f'{x:.1e}' for x in ea.tolist()])
