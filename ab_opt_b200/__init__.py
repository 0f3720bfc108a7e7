"""ab_opt_b200: B200-native (sm_100a) implementation of ab_opt's reverse-diffusion sampling hot
path -- FullDPM.sample / optimize -> EpsilonNet -> GAEncoder (invariant point attention) -> SO(3) /
R^3 / categorical transitions -- as hand-written CUDA kernels behind a C ABI
(include/abopt_b200.h, ab_opt_b200/_lib/libabopt_b200.so) and a thin PyTorch-facing mirror of the
reference's module interface.  There is no CPU fallback.
"""
from . import _capi
from ._capi import AboptError, launch_count
from .modules.encoders.ga import GABlock, GAEncoder
from .modules.encoders.pair import PairEmbedding, ResidueEmbedding
from . import post
from .post import reconstruct_backbone_partially, calc_per_rmsd, calc_avg_rmsd, rank_commoness
from .modules.diffusion.dpm_full import EpsilonNet, FullDPM, FullDPMAbDesign
from .models.diffab import DiffusionAntibodyDesign
from .modules.diffusion.transition import (VarianceSchedule, PositionTransition, RotationTransition,
                                           AminoacidCategoricalTransition)

__all__ = ['GABlock', 'GAEncoder', 'PairEmbedding', 'ResidueEmbedding', 'reconstruct_backbone_partially', 'calc_per_rmsd', 'calc_avg_rmsd',
           'rank_commoness', 'EpsilonNet', 'FullDPM', 'FullDPMAbDesign', 'DiffusionAntibodyDesign', 'VarianceSchedule',
           'PositionTransition', 'RotationTransition', 'AminoacidCategoricalTransition', 'AboptError',
           'launch_count', 'install_into_reference']


def install_into_reference(package='src', fused=True, embeddings=True):
    """Swap the reference's hot-path classes for the B200 ones so that its entry points
    (dock_pdb.py / design_pdb.py, configs/*.yml, checkpoints) run unchanged.

    package = 'src' (AbDock) or 'diffab' (AbDesign); the reference package must be importable.
    Must be called before the reference's `models/diffab.py` is imported.
    embeddings=False keeps the reference's own PairEmbedding / ResidueEmbedding (torch autograd): the configuration for
    train.py -- FullDPM.forward differentiates on the sm_100a path and hands d res_feat / d pair_feat to autograd, the embeddings'
    own backward is not part of the library (and fused is then off: the fused class contains the B200 embeddings).
    fused=True additionally registers ab_opt_b200.DiffusionAntibodyDesign under the reference's model name 'diffab'
    (models/_base.py:4-13), so `get_model(cfg.model).sample(batch)` is one device-resident encode + sample call.
    """
    import importlib
    dpm = importlib.import_module(f'{package}.modules.diffusion.dpm_full')
    ga = importlib.import_module(f'{package}.modules.encoders.ga')
    dpm.FullDPM = FullDPM if package == 'src' else FullDPMAbDesign
    dpm.EpsilonNet = EpsilonNet
    ga.GAEncoder = GAEncoder
    ga.GABlock = GABlock
    if embeddings:
        importlib.import_module(f'{package}.modules.encoders.pair').PairEmbedding = PairEmbedding      # models/diffab.py:28
        importlib.import_module(f'{package}.modules.encoders.residue').ResidueEmbedding = ResidueEmbedding  # models/diffab.py:27
    else:
        fused = False
    # after the loop: geometry.reconstruct_backbone_partially with the reference's own ideal-backbone tables; the ranking helpers
    # live in a runner module that needs lmdb / BioPython, so they are rebound only where that module imports
    K = importlib.import_module(f'{package}.utils.protein.constants')
    post.set_backbone_tables(K.backbone_atom_coordinates_tensor, K.bb_oxygen_coordinate_tensor)
    importlib.import_module(f'{package}.modules.common.geometry').reconstruct_backbone_partially = reconstruct_backbone_partially
    try:
        runner = importlib.import_module(f'{package}.tools.runner.design_for_testset')
        runner.calc_per_rmsd, runner.calc_avg_rmsd, runner.rank_commoness = calc_per_rmsd, calc_avg_rmsd, rank_commoness
    except ImportError:
        pass
    if fused:
        try:
            importlib.import_module(f'{package}.models')          # registers the reference's own class first ...
            base = importlib.import_module(f'{package}.models._base')
            flavour = 'abdock' if package == 'src' else 'abdesign'
            base._MODEL_DICT['diffab'] = lambda cfg: DiffusionAntibodyDesign(cfg, flavour=flavour)      # ... then ours replaces it
        except ImportError:
            pass
    return dpm, ga
