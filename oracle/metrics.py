"""Oracle (test infrastructure): CPU restatement of the reference's 3-D error metrics.

PINNED by tests/golden/metrics_golden.npz, produced by the real utils/eval_utils.py functions in the build container
(tests/golden/make_golden_metrics.py).  numpy, like the reference.
"""
import numpy as np


def scale_and_translation_transform_batch(P, T):
    """utils/eval_utils.py:105-125."""
    P_mean = np.mean(P, axis=-2, keepdims=True)
    P_trans = P - P_mean
    P_scale = np.sqrt(np.sum(P_trans ** 2, axis=(-2, -1), keepdims=True) / P.shape[-2])
    P_normalised = P_trans / P_scale
    T_mean = np.mean(T, axis=-2, keepdims=True)
    T_scale = np.sqrt(np.sum((T - T_mean) ** 2, axis=(-2, -1), keepdims=True) / T.shape[-2])
    return P_normalised * T_scale + T_mean


def procrustes_analysis_batch(S1, S2):
    """utils/eval_utils.py:62-102 (similarity transform of S1 onto S2, det(R) = +1)."""
    batch_size = S1.shape[0]
    S1 = S1.transpose(0, 2, 1)
    S2 = S2.transpose(0, 2, 1)
    mu1 = S1.mean(axis=2, keepdims=True)
    mu2 = S2.mean(axis=2, keepdims=True)
    X1 = S1 - mu1
    X2 = S2 - mu2
    var1 = (X1 ** 2).sum(axis=(1, 2))
    K = np.matmul(X1, X2.transpose(0, 2, 1))
    U, s, Vh = np.linalg.svd(K)
    V = Vh.transpose(0, 2, 1)
    Z = np.tile(np.eye(U.shape[1])[None, :, :], (batch_size, 1, 1))
    Z[:, -1, -1] *= np.sign(np.linalg.det(np.matmul(U, Vh)))
    R = np.matmul(np.matmul(V, Z), U.transpose(0, 2, 1))
    trace = np.matmul(R, K).diagonal(offset=0, axis1=-1, axis2=-2).sum(axis=-1)
    scale = (trace / var1)[..., None, None]
    t = mu2 - scale * np.matmul(R, mu1)
    S1_hat = scale * np.matmul(R, S1) + t
    return S1_hat.transpose(0, 2, 1)


def pointset_errors(pred, target):
    """pred (B,N,P,3), target (B,P,3) numpy -> dict of (B,N): the per-sample means that
    metrics/eval_metrics_tracker.py:119-280 sums (np.linalg.norm(..., axis=-1) then mean over the points)."""
    B, N, P, _ = pred.shape
    tgt = np.tile(target[:, None], (1, N, 1, 1))
    plain = np.linalg.norm(pred - tgt, axis=-1).mean(-1)
    sc = np.linalg.norm(scale_and_translation_transform_batch(pred, tgt) - tgt, axis=-1).mean(-1)
    pa = procrustes_analysis_batch(pred.reshape(B * N, P, 3), tgt.reshape(B * N, P, 3)).reshape(B, N, P, 3)
    pa = np.linalg.norm(pa - tgt, axis=-1).mean(-1)
    return {'plain': plain, 'sc': sc, 'pa': pa}
