"""Oracle: rigid-frame and SO(3) helpers (test infrastructure only, see oracle/__init__.py).

Citations are to /root/reference/AbDock/src/modules/common/{geometry,so3}.py
(the AbDesign mirror is identical except for its broken local_to_global, SURVEY.md finding 2).
"""
import math

import torch


def frame_to_global(R, t, p):
    """q = R p + t for points p (..., 3) attached to residues.  geometry.py:72-91.

    R (N,L,3,3), t (N,L,3), p (N,L,*,3) -> (N,L,*,3).
    """
    shape = p.shape
    N, L = shape[0], shape[1]
    pts = p.reshape(N, L, -1, 3)
    q = torch.einsum('nlab,nlkb->nlka', R, pts) + t[:, :, None, :]
    return q.reshape(shape)


def frame_to_local(R, t, q):
    """p = R^T (q - t).  geometry.py:94-113."""
    shape = q.shape
    N, L = shape[0], shape[1]
    pts = q.reshape(N, L, -1, 3) - t[:, :, None, :]
    p = torch.einsum('nlba,nlkb->nlka', R, pts)
    return p.reshape(shape)


def rotate_vector(R, p):
    """R p (no translation).  geometry.py:116-117."""
    return frame_to_global(R, torch.zeros_like(p), p)


def unit_vector(v, dim, eps=1e-6):
    """v / (|v| + eps).  geometry.py:32-33."""
    return v / (torch.linalg.norm(v, ord=2, dim=dim, keepdim=True) + eps)


def quat_1ijk_to_rotation(q):
    """(1 + b i + c j + d k), normalised, -> rotation matrix.  geometry.py:215-233."""
    b, c, d = q.unbind(-1)
    s = torch.sqrt(1 + b * b + c * c + d * d)
    a, b, c, d = 1 / s, b / s, c / s, d / s
    rows = [
        a * a + b * b - c * c - d * d, 2 * b * c - 2 * a * d, 2 * b * d + 2 * a * c,
        2 * b * c + 2 * a * d, a * a - b * b + c * c - d * d, 2 * c * d - 2 * a * b,
        2 * b * d - 2 * a * c, 2 * c * d + 2 * a * b, a * a - b * b - c * c + d * d,
    ]
    return torch.stack(rows, -1).reshape(q.shape[:-1] + (3, 3))


def quat_to_rotation(q):
    """Real-first quaternion (..., 4) -> rotation; normalises first.  geometry.py:148-174."""
    q = torch.nn.functional.normalize(q, dim=-1)
    r, i, j, k = q.unbind(-1)
    two_s = 2.0 / (q * q).sum(-1)
    rows = [
        1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
        two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
        two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j),
    ]
    return torch.stack(rows, -1).reshape(q.shape[:-1] + (3, 3))


def so3_exp(w):
    """Rodrigues exponential with the reference's +1e-8 / +2e-8 guards.  so3.py:33-57.

    The skew matrix convention is the reference's (so3.py:33-41):
        S = [[0, z, -y], [-z, 0, x], [y, -x, 0]]   for w = (x, y, z).
    """
    x, y, z = w.unbind(-1)
    o = torch.zeros_like(x)
    S = torch.stack([o, z, -y, -z, o, x, y, -x, o], -1).reshape(w.shape[:-1] + (3, 3))
    # so3.py:45 recovers the angle from S via skewsym_to_so3vec, i.e. |(S12, S20, S01)| = |w|
    ang = torch.linalg.norm(torch.stack([S[..., 1, 2], S[..., 2, 0], S[..., 0, 1]], -1), dim=-1)
    eye = torch.eye(3, dtype=w.dtype, device=w.device)
    b = (torch.sin(ang) + 1e-8) / (ang + 1e-8)
    c = (1 - torch.cos(ang) + 1e-8) / (ang * ang + 2e-8)
    return eye + b[..., None, None] * S + c[..., None, None] * (S @ S)


GRAD_ENABLED_CLAMP = False      # see so3_log / grad_enabled_semantics


class grad_enabled_semantics:
    """Context manager: evaluate the oracle as the reference evaluates it while autograd is ENABLED (a training step).
    The only numerical difference is log_rotation's clamp (so3.py:12-17: cos >= -0.999 instead of -1), which moves every
    rotation within 0.045 rad of pi -- about 3 % of uniformly random rotations.  Sampling, optimize() and the validation loop
    run under torch.no_grad() and use -1; that is the oracle's default."""

    def __enter__(self):
        global GRAD_ENABLED_CLAMP
        self.prev, GRAD_ENABLED_CLAMP = GRAD_ENABLED_CLAMP, True

    def __exit__(self, *exc):
        global GRAD_ENABLED_CLAMP
        GRAD_ENABLED_CLAMP = self.prev


def so3_log(R, grad_mode=None):
    """Log map to the so(3) vector, replicated op-for-op (ill-conditioned near pi on purpose).

    so3.py:10-30,60-63.  `grad_mode` selects the -0.999 clamp the reference uses when autograd
    is enabled (training); sampling runs under no_grad -> clamp at -1.  None -> the module switch
    (`grad_enabled_semantics`), which is off by default.
    """
    if grad_mode is None:
        grad_mode = GRAD_ENABLED_CLAMP
    tr = R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2]
    lo = -0.999 if grad_mode else -1.0
    cos_t = ((tr - 1) / 2).clamp_min(lo)
    sin_t = torch.sqrt(1 - cos_t * cos_t)
    theta = torch.acos(cos_t)
    coef = (theta + 1e-8) / (2 * sin_t + 2e-8)
    A = coef[..., None, None] * (R - R.transpose(-1, -2))
    return torch.stack([A[..., 1, 2], A[..., 2, 0], A[..., 0, 1]], -1)


def uniform_so3_from_gauss4(g4):
    """random_uniform_so3 given its N(0,1) draw g4 (..., 4).  so3.py:66-68."""
    q = torch.nn.functional.normalize(g4, dim=-1)
    return so3_log(quat_to_rotation(q))


PI = math.pi
