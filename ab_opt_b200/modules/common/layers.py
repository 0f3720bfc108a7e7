"""Parameter containers whose names match the reference (common/layers.py:108-160)."""
import torch
import torch.nn as nn


class LayerNorm(nn.Module):
    """Holds `gamma` / `beta` (NOT weight / bias) like the reference's hand-written LayerNorm
    (common/layers.py:108-160).  The arithmetic itself runs inside the fused CUDA kernels."""

    def __init__(self, normal_shape, gamma=True, beta=True, epsilon=1e-10):
        super().__init__()
        n = normal_shape if isinstance(normal_shape, int) else normal_shape[-1]
        self.normal_shape = torch.Size((n,))
        self.epsilon = epsilon
        if epsilon != 1e-10:
            raise ValueError('the CUDA kernels implement the reference default epsilon=1e-10 only')
        self.gamma = nn.Parameter(torch.ones(n)) if gamma else None
        self.beta = nn.Parameter(torch.zeros(n)) if beta else None
        if not (gamma and beta):
            raise ValueError('gamma and beta are both required by the CUDA kernels')
