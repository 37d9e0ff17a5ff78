"""Oracle (test infrastructure): CPU restatement of the reference's 3-D error metrics.

PINNED by tests/golden/metrics_golden.npz, produced by the real utils/eval_utils.py functions in the build container
(tests/golden/make_golden_metrics.py).  numpy float64 internally, written from the definitions:

  scale + translation correction (utils/eval_utils.py:105-125): centre the prediction, rescale it so that its RMS
      distance from its centroid equals the target's, move it to the target's centroid;
  Procrustes alignment (utils/eval_utils.py:62-102): the similarity transform (s, R, t), det R = +1, minimising
      sum_i || s R p_i + t - q_i ||^2  (Umeyama / Kabsch with the reflection fix on the smallest singular direction);
  error = mean over the points of the Euclidean distance (metrics/eval_metrics_tracker.py:119-280).
"""
import numpy as np


def _centroid(x):
    return x.mean(axis=-2, keepdims=True)


def align_scale_translation(pred, target):
    """(..., P, 3) x2 -> pred moved onto the target's centroid and RMS radius (utils/eval_utils.py:105-125)."""
    pc, tc = pred - _centroid(pred), target - _centroid(target)
    radius = lambda c: np.sqrt((c * c).sum(axis=(-2, -1), keepdims=True) / c.shape[-2])
    return pc * (radius(tc) / radius(pc)) + _centroid(target)


def align_procrustes(pred, target):
    """(n, P, 3) x2 -> s R pred + t with the optimal similarity transform per set (utils/eval_utils.py:62-102)."""
    mu_p, mu_t = _centroid(pred), _centroid(target)
    pc, tc = pred - mu_p, target - mu_t
    cov = np.einsum('npi,npj->nij', pc, tc)                       # sum_p pc_p tc_p^T  (3x3 per set)
    u, sing, vt = np.linalg.svd(cov)
    flip = np.sign(np.linalg.det(u @ vt))                         # -1 where the best orthogonal map is a reflection
    fix = np.ones_like(sing)
    fix[:, -1] = flip
    rot = np.einsum('nji,nj,nkj->nik', vt, fix, u)                # V diag(fix) U^T
    scale = (sing * fix).sum(-1) / (pc * pc).sum(axis=(-2, -1))   # trace(R cov) / ||pc||^2
    return scale[:, None, None] * np.einsum('nij,npj->npi', rot, pc) + mu_t


def pointset_errors(pred, target):
    """pred (B,N,P,3), target (B,P,3) -> dict of (B,N) mean point errors: plain, scale-corrected, Procrustes-aligned."""
    B, N, P, _ = pred.shape
    p = pred.astype(np.float64)
    t = np.broadcast_to(target.astype(np.float64)[:, None], p.shape)
    dist = lambda a: np.linalg.norm(a - t, axis=-1).mean(-1)
    pa = align_procrustes(p.reshape(B * N, P, 3), t.reshape(B * N, P, 3)).reshape(p.shape)
    return {'plain': dist(p), 'sc': dist(align_scale_translation(p, t)), 'pa': dist(pa)}


def sample_stats(points, target=None, weights=None):
    """points (B,N,P,D), target (B,P,D) or None, weights (B,P) or None -> per-frame sample diversity and samples-L2E
    (metrics/eval_metrics_tracker.py:397-433 and :339-374, the per-frame values those blocks append)."""
    x = points.astype(np.float64)
    w = np.ones(x.shape[0:1] + x.shape[2:3]) if weights is None else weights.astype(np.float64)
    dist_from_mean = np.linalg.norm(x - x.mean(axis=1)[:, None], axis=-1) * w[:, None, :]
    out = {'diversity': dist_from_mean.mean(axis=(1, 2))}
    if target is not None:
        l2e = np.linalg.norm(x - target.astype(np.float64)[:, None], axis=-1) * w[:, None, :]
        out['l2e'] = l2e.sum(axis=(1, 2)) / (w.sum(axis=-1) * x.shape[1])
    return out
