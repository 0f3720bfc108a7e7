"""IGSO(3) angle tables with the reference's buffer names (common/so3.py:70-109).

Only construction (an init-time CPU precompute, exactly as in the reference) and the buffers
live here; sampling happens in the transition kernels."""
import math

import torch
import torch.nn as nn


class ApproxAngularDistribution(nn.Module):

    def __init__(self, stddevs, std_threshold=0.1, num_bins=8192, num_iters=1024):
        super().__init__()
        if num_bins != 8192:
            raise ValueError('the CUDA kernels assume the reference default of 8192 angle bins')
        self.std_threshold, self.num_bins, self.num_iters = std_threshold, num_bins, num_iters
        self.register_buffer('stddevs', torch.FloatTensor(stddevs))
        self.register_buffer('approx_flag', self.stddevs <= std_threshold)
        # density of the angle of IGSO(3)(sigma), truncated series, on a uniform grid over [0, pi]
        x = torch.linspace(0, math.pi, num_bins)
        l = torch.arange(0, num_iters)[None, :]
        lead = ((1 - torch.cos(x)) / math.pi)[:, None]
        ratio = (torch.sin((l + 0.5) * x[:, None]) + 1e-6) / (torch.sin(x[:, None] / 2) + 1e-6)
        rows = []
        for e in self.stddevs.tolist():
            a = (2 * l + 1) * torch.exp(-l * (l + 1) * (e ** 2))
            rows.append(torch.nan_to_num((lead * a * ratio).sum(dim=1)).clamp_min(0))
        self.register_buffer('X', x[None].repeat(len(rows), 1))
        self.register_buffer('Y', torch.stack(rows, dim=0))
