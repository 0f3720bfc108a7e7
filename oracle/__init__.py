"""CPU oracle for the ab_opt denoising hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, in plain PyTorch on the CPU (fp32 or fp64), the algorithm of the
reference path  FullDPM.sample -> EpsilonNet -> GAEncoder (6x invariant point attention)
-> SO(3) / R^3 / categorical transitions  (reference: AbDock/src/modules/..., mirrored in
AbDesign/diffab/modules/...).  Every function cites the reference file:line it follows.

Rules:
  * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
    legs may import anything from here -- and only as the checker / CPU baseline.
  * The product (ab_opt_b200/) never imports this package; it fails loudly if the CUDA
    library is missing.  There is no CPU fallback.

Parity pinning: the reference ships no golden vectors or tests, so
the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, run in the build container
by tests/golden/make_golden.py (imports /root/reference/AbDock/src read-only) and
committed as fixtures under tests/golden/*.npz; tests/test_oracle_golden.py replays them
anywhere, and tests/test_oracle_vs_reference.py compares live when /root/reference exists.
"""
from . import geometry, ipa, epsnet, transitions, sampler, weights  # noqa: F401
# training.py (FullDPM.forward, losses + autograd step), pair_embed.py (PairEmbedding / ResidueEmbedding, the step before the
# loop) and post.py (backbone reconstruction, RMSD ranking, the steps after it) are imported where they are used.
