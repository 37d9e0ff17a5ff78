"""Oracle (test infrastructure): rotation helpers of the hot path, restated on CPU torch.

Follows /root/reference/utils/rigid_transform_utils.py (reference-owned, PINNED by
tests/golden/so3_golden.npz) and smplx 0.1.26 ``lbs.batch_rodrigues`` ([upstream], unpinned).
Every function cites the reference lines it restates.  dtype behaviour is kept: the exp/log maps
run in float64 exactly where the reference asserts float64.
"""
import itertools
import math

import torch


def rot6d_to_rotmat(x):
    """utils/rigid_transform_utils.py:86-100.  (B,6) row-major [R11,R12,R21,R22,R31,R32] -> (B,3,3);
    Gram-Schmidt on the two columns, third column = cross product, columns stacked on the last dim.
    F.normalize semantics: v / max(||v||, 1e-12)."""
    x = x.reshape(-1, 3, 2)
    a1, a2 = x[:, :, 0], x[:, :, 1]
    b1 = a1 / a1.norm(dim=1, keepdim=True).clamp_min(1e-12)
    u2 = a2 - (b1 * a2).sum(dim=1, keepdim=True) * b1
    b2 = u2 / u2.norm(dim=1, keepdim=True).clamp_min(1e-12)
    b3 = torch.linalg.cross(b1, b2, dim=1)
    return torch.stack((b1, b2, b3), dim=-1)


def rotmat_to_rot6d(R):
    """utils/rigid_transform_utils.py:103-116 with stack_columns=False: first two columns, row-major."""
    return R[:, :, :2].reshape(-1, 6)


def hat(v):
    """utils/rigid_transform_utils.py:142-163.  (...,3) -> skew-symmetric (...,3,3)."""
    z = torch.zeros_like(v[..., 0])
    rows = [torch.stack((z, -v[..., 2], v[..., 1]), -1),
            torch.stack((v[..., 2], z, -v[..., 0]), -1),
            torch.stack((-v[..., 1], v[..., 0], z), -1)]
    return torch.stack(rows, -2)


def vee(m):
    """utils/rigid_transform_utils.py:166-179."""
    return torch.stack((-m[..., 1, 2], m[..., 0, 2], -m[..., 0, 1]), -1)


def so3_exp(v):
    """utils/rigid_transform_utils.py:182-201.  float64 Rodrigues with Taylor guard at theta<=1e-10."""
    assert v.dtype == torch.float64
    theta = v.norm(dim=-1)
    big = theta > 1e-10
    th = torch.where(big, theta, torch.ones_like(theta))
    alpha = torch.where(big, torch.sin(th) / th, 1 - th ** 2 / 6)
    beta = torch.where(big, (1 - torch.cos(th)) / th ** 2, 0.5 - th ** 2 / 24)
    K = hat(v)
    eye = torch.eye(3, dtype=v.dtype)
    return eye + alpha[..., None, None] * K + beta[..., None, None] * (K @ K)


def so3_log_pi(r, theta):
    """utils/rigid_transform_utils.py:240-279.  Near-pi branch: axis magnitudes from the symmetric
    part, sign pattern chosen among the 8 candidates (order = itertools.product([0,1],repeat=3)*2-1,
    first minimum wins) by ||R - exp(x)||_F^2.  r: (M,3,3) f64, theta: (M,1,1)."""
    sym = 0.5 * (r + r.transpose(-1, -2))
    eye = torch.eye(3, dtype=r.dtype).expand_as(sym)
    z = theta ** 2 / (1 - torch.cos(theta)) * (sym - eye)
    q1, q2, q3 = z[..., 0, 0], z[..., 1, 1], z[..., 2, 2]
    x1 = torch.sqrt(torch.clamp(q1 - q2 - q3, min=1e-8) / 2)
    x2 = torch.sqrt(torch.clamp(-q1 + q2 - q3, min=1e-8) / 2)
    x3 = torch.sqrt(torch.clamp(-q1 - q2 + q3, min=1e-8) / 2)
    x = torch.stack([x1, x2, x3], -1).reshape(-1, 3)
    r = r.reshape(-1, 3, 3)
    signs = torch.tensor(list(itertools.product([0, 1], repeat=3)), dtype=x.dtype) * 2 - 1
    cand = signs.view(8, 1, 3) * x[None]
    diff = (r[None] - so3_exp(cand)).pow(2).sum(-1).sum(-1)
    sel = torch.argmin(diff, dim=0)
    return cand[sel, torch.arange(len(sel))]


def so3_log(r):
    """utils/rigid_transform_utils.py:204-237, returned as the axis-angle vector (so3_vee of the log).
    theta = acos(clamp((tr-1)/2)); theta/sin(theta) * skew part; Taylor for theta<1e-20;
    |pi-theta|<1e-2 handled by so3_log_pi."""
    assert r.dtype == torch.float64
    shape = r.shape[:-2]
    r = r.reshape(-1, 3, 3)
    anti = 0.5 * (r - r.transpose(-1, -2))
    cos_t = (0.5 * (r[:, 0, 0] + r[:, 1, 1] + r[:, 2, 2] - 1)).clamp(-1, 1)
    theta = torch.acos(cos_t)
    ratio = theta / torch.sin(theta)
    ratio = torch.where(theta < 1e-20, 1 + theta ** 2 / 6, ratio)
    x = vee(ratio[:, None, None] * anti)
    near_pi = (math.pi - theta).abs() < 1e-2
    if near_pi.any():
        idx = near_pi.nonzero()[:, 0]
        x[idx] = so3_log_pi(r[idx], theta[idx, None, None])
    return x.reshape(*shape, 3)


def so3_xset(x):
    """utils/rigid_transform_utils.py:282-295 with k_max=1: the two other pre-images of exp,
    order k = [-1, +1].  Returns (2, ..., 3).  ||x||==0 gives NaN exactly as the reference does
    (masked away by the caller)."""
    n = x.norm(dim=-1, keepdim=True)
    k = torch.tensor([-1.0, 1.0], dtype=x.dtype).view(2, *([1] * x.dim()))
    return x[None] / n[None] * (n[None] + 2 * math.pi * k)


def so3_log_abs_det_jacobian(x):
    """utils/rigid_transform_utils.py:298-314.  log((2-2cos t)/t^2) in float64, cast to x.dtype."""
    n = x.double().norm(dim=-1)
    big = n > 1e-10
    n1 = torch.where(big, n, torch.ones_like(n))
    ratio = torch.where(big, (2 - 2 * torch.cos(n1)) / n1 ** 2, 1 - n1 ** 2 / 12)
    return torch.log(ratio).to(x.dtype)


def batch_rodrigues(rot_vecs):
    """[upstream, unpinned] smplx 0.1.26 lbs.batch_rodrigues, called at
    models/humaniflow_model.py:299 (fp32 point estimate) and inside smplx lbs when pose2rot=True.
    angle = ||v + 1e-8||, axis = v / angle, R = I + sin*K + (1-cos)*K@K, all in the input dtype."""
    angle = torch.norm(rot_vecs + 1e-8, dim=1, keepdim=True)
    axis = rot_vecs / angle
    K = hat(axis)
    cos = torch.cos(angle)[:, :, None]
    sin = torch.sin(angle)[:, :, None]
    eye = torch.eye(3, dtype=rot_vecs.dtype)[None]
    return eye + sin * K + (1 - cos) * torch.bmm(K, K)
