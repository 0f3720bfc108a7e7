"""Helpers to import the UNMODIFIED reference (read-only, build container only).

Used by tests/golden/make_golden.py and tests/test_oracle_vs_reference.py; nothing here is
reachable from the GPU tests, smoke() or bench.py (/root/reference does not exist on the GPU box).
"""
import importlib
import os
import sys

REF_ROOT = os.environ.get('ABOPT_REFERENCE', '/root/reference')
ABDOCK = os.path.join(REF_ROOT, 'AbDock')


def reference_available():
    return os.path.isdir(os.path.join(ABDOCK, 'src', 'modules', 'diffusion'))


def import_abdock():
    """Returns the reference modules of the AbDock flavour (imports `src.*` from AbDock/)."""
    if ABDOCK not in sys.path:
        sys.path.insert(0, ABDOCK)
    mods = {}
    for name in ('src.modules.diffusion.dpm_full', 'src.modules.diffusion.transition',
                 'src.modules.encoders.ga', 'src.modules.common.so3', 'src.modules.common.geometry',
                 'src.modules.common.layers'):
        mods[name.rsplit('.', 1)[1]] = importlib.import_module(name)
    return mods


def build_reference_fulldpm(W, num_layers=6, obj='pred_x0', num_bins=40):
    """Reference FullDPM with state-dict `W` loaded STRICTLY (proves key/shape compatibility)."""
    import torch
    m = import_abdock()
    # the histogram precompute in __init__ is slow (2 x 101 sigmas); shrink it, then load real buffers
    tr_opt = dict(angular_distrib_fwd_opt=dict(num_iters=2), angular_distrib_inv_opt=dict(num_iters=2))
    model = m['dpm_full'].FullDPM(128, 64, num_steps=100, eps_net_opt=dict(num_layers=num_layers),
                                  trans_rot_opt=tr_opt, obj=obj, num_bins=num_bins)
    missing = model.load_state_dict(W, strict=True)
    model.eval()
    return model, m


def load_reference_functions(rel_path, names):
    """Compile the named top-level functions of a reference source file WITHOUT importing the module (the runner modules pull
    in lmdb / Bio / pyrosetta, which are absent here).  The function bodies are the reference's, unmodified."""
    import ast
    import torch
    path = os.path.join(ABDOCK, rel_path)
    tree = ast.parse(open(path).read())
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert len(keep) == len(names), f'{names} not all found in {path}'
    ns = {'torch': torch}
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, 'exec'), ns)
    return {n: ns[n] for n in names}
