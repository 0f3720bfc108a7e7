"""Live comparison oracle <-> UNMODIFIED reference.  Build container only: skipped wherever
/root/reference is absent (e.g. the GPU box).  Includes a trained checkpoint so parity is
exercised at realistic activations."""
import os
import pickle
import sys
import types

import pytest
import torch

from refload import reference_available, build_reference_fulldpm, REF_ROOT
from oracle import weights, epsnet, sampler, geometry as G

pytestmark = pytest.mark.skipif(not reference_available(), reason='reference tree not present')


def _load_ckpt_state(path):
    """torch.load needs easydict / dynamic_yaml to unpickle the config; stub them."""
    for name in ('easydict', 'dynamic_yaml', 'dynamic_yaml.yaml_wrappers'):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    class _Dict(dict):
        def __setstate__(self, st):
            pass
    class _Seq(list):
        def __setstate__(self, st):
            pass
    sys.modules['easydict'].EasyDict = _Dict
    yw = sys.modules['dynamic_yaml.yaml_wrappers']
    yw.YamlDict, yw.YamlList = _Dict, _Seq
    ck = torch.load(path, map_location='cpu', weights_only=False)
    return {k[len('diffusion.'):]: v for k, v in ck['model'].items() if k.startswith('diffusion.')}


@torch.no_grad()
def test_trained_checkpoint_eps_net_and_optimize():
    path = os.path.join(REF_ROOT, 'AbDock/reproduction/dock_single_cdr/250000.pt')
    if not os.path.exists(path):
        pytest.skip('checkpoint missing')
    W = _load_ckpt_state(path)
    model, m = build_reference_fulldpm(W, num_layers=6, obj='pred_x0')
    inp = weights.synthetic_inputs(5, 2, 40, gen_slices=((16, 26),), ragged=True)
    beta = W['trans_pos.var_sched.betas'][80].expand(2)
    ref = model.eps_net(inp['v'], inp['p'] / 10, inp['s'], inp['res_feat'], inp['pair_feat'], beta,
                        inp['mask_generate'], inp['mask_res'])
    out = epsnet.eps_net(W, inp['v'], inp['p'] / 10, inp['s'], inp['res_feat'], inp['pair_feat'], beta,
                         inp['mask_generate'], inp['mask_res'])
    # Arbiter = fp64 evaluation of the oracle: the fp32 oracle must be as close to it as the
    # fp32 reference is (trained weights amplify fp32 round-off to ~5e-5 on R_next).
    W64 = weights.cast(W, torch.double)
    i64 = {k: (v.double() if v.is_floating_point() else v) for k, v in inp.items()}
    o64 = epsnet.eps_net(W64, i64['v'], i64['p'] / 10, i64['s'], i64['res_feat'], i64['pair_feat'],
                         beta.double(), i64['mask_generate'], i64['mask_res'])
    for i, floor in ((1, 1e-6), (2, 1e-6), (3, 1e-7), (4, 1e-6)):
        e_ref = (ref[i] - o64[i]).abs().max().item()
        e_orc = (out[i] - o64[i]).abs().max().item()
        assert e_orc <= 2 * e_ref + floor, (i, e_orc, e_ref)
        assert e_ref < 1e-3

    # optimize(): forward-noise to step 3, then 3 reverse steps, same seed on both sides
    torch.manual_seed(99)
    tr = model.optimize(inp['v'], inp['p'], inp['s'], 3, inp['res_feat'], inp['pair_feat'],
                        inp['mask_generate'], inp['mask_res'])
    to = sampler.sample(W, inp['v'], inp['p'], inp['s'], inp['res_feat'], inp['pair_feat'], inp['mask_generate'],
                        inp['mask_res'], obj='pred_x0', gen=torch.Generator().manual_seed(99), start_step=3)
    for t in (3, 2):
        assert torch.equal(to[t][2], tr[t][2])
        torch.testing.assert_close(to[t][1], tr[t][1], rtol=1e-4, atol=2e-3)
        torch.testing.assert_close(G.so3_exp(to[t][0]), G.so3_exp(tr[t][0]), rtol=0, atol=1e-3)
    torch.testing.assert_close(to[2][3], tr[2][3], rtol=1e-4, atol=1e-4)      # pRMSD
    torch.testing.assert_close(to[2][4], tr[2][4], rtol=1e-4, atol=1e-5)      # perplexity (no mask in optimize)


@torch.no_grad()
def test_abdesign_flavour_eps_net():
    """AbDesign's EpsilonNet (no pRMSD head) with AbDock's working local_to_global patched in
    (the AbDesign copy of local_to_global raises as shipped, SURVEY.md finding 2)."""
    abdesign = os.path.join(REF_ROOT, 'AbDesign')
    if not os.path.isdir(abdesign):
        pytest.skip('AbDesign tree missing')
    # diffab.utils.misc drags in BioPython / easydict / torch_scatter, none of which the hot path
    # uses: transition.py only imports four helper names from it.  Stub that one module.
    misc = types.ModuleType('diffab.utils.misc')
    for attr in ('hotspot_distance_fn', 'pair2edge', 'batchfy', 'clash_loss'):
        setattr(misc, attr, None)
    sys.modules.setdefault('diffab.utils.misc', misc)
    sys.path.insert(0, abdesign)
    try:
        import importlib
        geo = importlib.import_module('diffab.modules.common.geometry')
        from refload import import_abdock
        geo.local_to_global = import_abdock()['geometry'].local_to_global
        dpm = importlib.import_module('diffab.modules.diffusion.dpm_full')
    except Exception as e:                                   # missing third-party deps of the AbDesign tree
        pytest.skip(f'AbDesign flavour not importable here: {type(e).__name__}: {e}')
    finally:
        sys.path.remove(abdesign)
    W = weights.make_state_dict(seed=3, num_layers=2, flavour='abdesign')
    net = dpm.EpsilonNet(128, 64, num_layers=2)
    net.load_state_dict({k[len('eps_net.'):]: v for k, v in W.items() if k.startswith('eps_net.')}, strict=True)
    inp = weights.synthetic_inputs(6, 2, 20, gen_slices=((4, 10),))
    beta = W['trans_pos.var_sched.betas'][30].expand(2)
    ref = net(inp['v'], inp['p'] / 10, inp['s'], inp['res_feat'], inp['pair_feat'], beta, inp['mask_generate'],
              inp['mask_res'])
    out = epsnet.eps_net(W, inp['v'], inp['p'] / 10, inp['s'], inp['res_feat'], inp['pair_feat'], beta,
                         inp['mask_generate'], inp['mask_res'])
    assert len(ref) == 4 and len(out) == 4
    torch.testing.assert_close(out[1], ref[1], rtol=0, atol=5e-6)
    torch.testing.assert_close(out[2], ref[2], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(out[3], ref[3], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize('obj', ['pred_x0', 'pred_noise'])
def test_training_forward_losses(obj):
    """FullDPM.forward (dpm_full.py:156-234) vs oracle.training.loss_forward on replayed draws: the reference draws from
    the default generator inside its three add_noise calls, the oracle receives the same stream as a noise record."""
    from oracle import training, transitions as T
    W = weights.make_state_dict(seed=5, num_layers=2, flavour='abdock')
    model, _ = build_reference_fulldpm(W, num_layers=2, obj=obj)
    inp = weights.synthetic_inputs(8, 3, 24, gen_slices=((0, 6), (14, 18)), ragged=True)
    t = torch.tensor([57, 3, 99])
    torch.manual_seed(11)
    with torch.no_grad():
        ref = model(inp['v'], inp['p'], inp['s'], inp['res_feat'], inp['pair_feat'], inp['mask_generate'], inp['mask_res'],
                    denoise_structure=True, denoise_sequence=True, t=t)
    noise = T.draw_step_noise(3, 24, torch.Generator().manual_seed(11))
    got = training.loss_forward(W, inp['v'], inp['p'], inp['s'], inp['res_feat'], inp['pair_feat'], inp['mask_generate'],
                                inp['mask_res'], True, True, t, noise, flavour='abdock', obj=obj)
    assert sorted(got) == sorted(ref)
    for k in ref:
        torch.testing.assert_close(got[k], ref[k], rtol=2e-5, atol=2e-6, msg=lambda m, k=k: f'{k}: {m}')


@torch.no_grad()
@pytest.mark.parametrize('ckpt,A', [('dock_single_cdr/250000.pt', 15), ('seq_design/300000.pt', 5)])
def test_pair_embedding_trained_weights(ckpt, A):
    """oracle.pair_embed vs the reference PairEmbedding (encoders/pair.py:37-101) with the shipped `pair_embed.*` weights."""
    path = os.path.join(REF_ROOT, 'AbDock/reproduction', ckpt)
    if not os.path.exists(path):
        pytest.skip('checkpoint missing')
    if os.path.join(REF_ROOT, 'AbDock') not in sys.path:
        sys.path.insert(0, os.path.join(REF_ROOT, 'AbDock'))
    from src.modules.encoders.pair import PairEmbedding
    from oracle import pair_embed as PE
    _load_ckpt_state(path)                      # installs the unpickling stubs
    ck = torch.load(path, map_location='cpu', weights_only=False)
    W = {k[len('pair_embed.'):]: v for k, v in ck['model'].items() if k.startswith('pair_embed.')}
    assert W['aapair_to_distcoef.weight'].shape[1] == A * A
    assert set(W) == set(PE.make_state_dict(0, A)) and all(W[k].shape == v.shape for k, v in PE.make_state_dict(0, A).items())
    ref = PairEmbedding(64, A)
    ref.load_state_dict(W, strict=True)
    ref.eval()
    inp = PE.synthetic_complex(9, 2, 32)
    m = inp['context_mask']
    args = (inp['aa'], inp['res_nb'], inp['chain_nb'], inp['pos_atoms'], inp['mask_atoms'])
    for sm, qm in ((None, None), (m, m)):
        want = ref(*args, structure_mask=sm, sequence_mask=qm)
        got = PE.pair_embedding(W, *args, sm, qm)
        off = ~torch.eye(32, dtype=torch.bool)[None, :, :, None]
        torch.testing.assert_close(got * off, want * off, rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(got, want, rtol=0, atol=5e-3)


@torch.no_grad()
@pytest.mark.parametrize('ckpt,A', [('dock_single_cdr/250000.pt', 15), ('seq_design/300000.pt', 5)])
def test_residue_embedding_trained_weights(ckpt, A):
    """oracle residue_embedding vs the reference ResidueEmbedding (encoders/residue.py:27-94) with the shipped `residue_embed.*` weights."""
    path = os.path.join(REF_ROOT, 'AbDock/reproduction', ckpt)
    if not os.path.exists(path):
        pytest.skip('checkpoint missing')
    if os.path.join(REF_ROOT, 'AbDock') not in sys.path:
        sys.path.insert(0, os.path.join(REF_ROOT, 'AbDock'))
    from src.modules.encoders.residue import ResidueEmbedding
    from oracle import pair_embed as PE
    _load_ckpt_state(path)
    ck = torch.load(path, map_location='cpu', weights_only=False)
    W = {k[len('residue_embed.'):]: v for k, v in ck['model'].items() if k.startswith('residue_embed.')}
    proto = PE.make_residue_state_dict(0, A)
    assert set(W) == set(proto) and all(W[k].shape == v.shape for k, v in proto.items())
    ref = ResidueEmbedding(128, A)
    ref.load_state_dict(W, strict=True)
    ref.eval()
    inp = PE.synthetic_complex(9, 2, 32)
    ft = torch.randint(0, 4, (2, 32), generator=torch.Generator().manual_seed(9))
    m = inp['context_mask']
    args = (inp['aa'], inp['res_nb'], inp['chain_nb'], inp['pos_atoms'], inp['mask_atoms'], ft)
    for sm, qm in ((None, None), (m, m), (m, None)):
        torch.testing.assert_close(PE.residue_embedding(W, *args, sm, qm), ref(*args, structure_mask=sm, sequence_mask=qm),
                                   rtol=1e-5, atol=1e-5)


@torch.no_grad()
def test_post_loop_functions():
    """oracle.post vs the reference's reconstruct_backbone_partially (geometry.py:450-480) and the ranking helpers of
    tools/runner/design_for_testset.py:556-589 (compiled from the source file: the module itself needs lmdb / BioPython)."""
    if os.path.join(REF_ROOT, 'AbDock') not in sys.path:
        sys.path.insert(0, os.path.join(REF_ROOT, 'AbDock'))
    from src.modules.common.geometry import reconstruct_backbone_partially
    from src.utils.protein import constants as K
    from refload import load_reference_functions
    from oracle import post, pair_embed as PE
    fn = load_reference_functions('src/tools/runner/design_for_testset.py', ['calc_per_rmsd', 'calc_avg_rmsd', 'rank_commoness'])
    inp = PE.synthetic_complex(12, 3, 50)
    g = torch.Generator().manual_seed(4)
    R, t = G.so3_exp(torch.randn(3, 50, 3, generator=g)), torch.randn(3, 50, 3, generator=g) * 20
    aa = torch.randint(0, 21, (3, 50), generator=g)
    for rec in (~inp['context_mask'], torch.ones(3, 50, dtype=torch.bool)):
        args = (inp['pos_atoms'], R, t, aa, inp['chain_nb'], inp['res_nb'], inp['mask_atoms'], rec)
        want = reconstruct_backbone_partially(*args)
        got = post.reconstruct_backbone_partially(*args, K.backbone_atom_coordinates_tensor, K.bb_oxygen_coordinate_tensor)
        torch.testing.assert_close(got[0], want[0], rtol=1e-6, atol=1e-5)
        assert torch.equal(got[1], want[1])
    S = torch.randn(40, 25, 3, generator=g) * 2 + torch.randn(1, 25, 3, generator=g) * 10
    torch.testing.assert_close(post.pairwise_rmsd(S), fn['calc_per_rmsd'](S), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(post.average_rmsd(S), fn['calc_avg_rmsd'](S), rtol=1e-6, atol=0)
    assert torch.equal(post.rank_commonness(S, 7), fn['rank_commoness'](S, 7))


class _Cfg(dict):
    """EasyDict stand-in (attribute access, nested): the reference's configs are EasyDicts (utils/misc.py load_config)."""
    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k)
        return _Cfg(v) if isinstance(v, dict) else v


def test_drop_in_through_the_reference_registry():
    """The boundary itself: after ab_opt_b200.install_into_reference('src'), the reference's own
    `get_model(cfg.model)` (models/_base.py:12-13, models/diffab.py:19-37) builds the B200 classes from the reference's YAML, and
    `load_state_dict(ckpt['model'])` (tools/runner/design_for_pdb.py:94) loads the shipped checkpoint STRICTLY into them.
    Runs in a subprocess so that the rebinding does not leak into the other tests of this file.  CPU only: nothing is computed."""
    import subprocess
    import textwrap
    path = os.path.join(REF_ROOT, 'AbDock/reproduction/dock_single_cdr/250000.pt')
    if not os.path.exists(path):
        pytest.skip('checkpoint missing')
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = textwrap.dedent(f'''
        import sys, yaml, torch
        sys.path[:0] = [{os.path.join(REF_ROOT, 'AbDock')!r}, {repo!r}, {os.path.join(repo, 'tests')!r}]
        import ab_opt_b200
        ab_opt_b200.install_into_reference('src')
        from src.models import get_model
        import src.modules.common.geometry as geo
        from test_oracle_vs_reference import _Cfg, _load_ckpt_state
        cfg = _Cfg(yaml.safe_load(open({os.path.join(REF_ROOT, 'AbDock/configs/train/dock_single.yml')!r})))
        model = get_model(cfg.model)
        assert type(model) is ab_opt_b200.DiffusionAntibodyDesign          # the fused encode + sample class, same state-dict keys
        assert type(model.diffusion) is ab_opt_b200.FullDPM and type(model.pair_embed) is ab_opt_b200.PairEmbedding
        assert type(model.residue_embed) is ab_opt_b200.ResidueEmbedding
        assert type(model.diffusion.eps_net.encoder) is ab_opt_b200.GAEncoder
        assert geo.reconstruct_backbone_partially is ab_opt_b200.reconstruct_backbone_partially
        _load_ckpt_state({path!r})
        ck = torch.load({path!r}, map_location='cpu', weights_only=False)
        res = model.load_state_dict(ck['model'], strict=True)
        assert not res.missing_keys and not res.unexpected_keys
        try:                                     # and there is no CPU path behind the classes
            model.pair_embed(*[torch.zeros(1, 2, dtype=torch.long)] * 3, torch.zeros(1, 2, 15, 3), torch.ones(1, 2, 15, dtype=torch.bool))
        except ab_opt_b200.AboptError as e:
            print('OK', len(ck['model']), 'tensors;', str(e)[:60])
    ''')
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.startswith('OK'), out.stdout + out.stderr[-3000:]
