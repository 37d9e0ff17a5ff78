"""Golden vectors for the 3-D error metrics from the REAL reference functions (utils/eval_utils.py) — build container only.
    python tests/golden/make_golden_metrics.py   ->  tests/golden/metrics_golden.npz"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, '/root/reference')
from utils.eval_utils import procrustes_analysis_batch, scale_and_translation_transform_batch  # noqa: E402

rs = np.random.RandomState(0)
B, N, P = 3, 4, 500
target = (rs.standard_normal((B, P, 3)) * 0.3).astype(np.float32)
# predictions = rotated / scaled / shifted / noisy copies of the target (one of them a reflection-prone near-planar case)
pred = np.zeros((B, N, P, 3), np.float32)
for b in range(B):
    for n in range(N):
        a = rs.standard_normal(3); a /= np.linalg.norm(a); ang = rs.uniform(0, 1.0)
        Kx = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
        R = np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx
        pred[b, n] = (rs.uniform(0.7, 1.3) * target[b] @ R.T + rs.standard_normal(3) * 0.2 + rs.standard_normal((P, 3)) * 0.02).astype(np.float32)
pred[0, 0, :, 2] *= 0.01          # nearly planar prediction
pred[1, 1] = target[1] * np.array([1, 1, -1], np.float32)   # mirror image: forces the det(R) = +1 correction
tgt = np.tile(target[:, None], (1, N, 1, 1))
sc = scale_and_translation_transform_batch(pred, tgt)
pa = procrustes_analysis_batch(pred.reshape(B * N, P, 3), tgt.reshape(B * N, P, 3)).reshape(B, N, P, 3)
np.savez(os.path.join(HERE, 'metrics_golden.npz'), pred=pred, target=target,
         plain=np.linalg.norm(pred - tgt, axis=-1).mean(-1), sc=np.linalg.norm(sc - tgt, axis=-1).mean(-1),
         pa=np.linalg.norm(pa - tgt, axis=-1).mean(-1), sc_points=sc[:, :, :5], pa_points=pa[:, :, :5])
print('wrote metrics_golden.npz')
