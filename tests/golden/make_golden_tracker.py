"""Golden vectors of the sample-diversity / samples-L2E per-frame metrics from the REAL reference class
metrics/eval_metrics_tracker.py::EvalMetricsTracker (build container only).
    python tests/golden/make_golden_tracker.py   ->  tests/golden/tracker_golden.npz
The tracker derives the input joints and their visibility from the model input's heatmap channels
(utils/label_conversions.py::convert_heatmaps_to_2Djoints_coordinates_torch); both are stored as fixtures."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, '/root/reference')
from metrics.eval_metrics_tracker import EvalMetricsTracker  # noqa: E402
from utils.label_conversions import convert_2Djoints_to_gaussian_heatmaps_torch, convert_heatmaps_to_2Djoints_coordinates_torch  # noqa: E402

rs = np.random.RandomState(3)
B, N, V = 3, 5, 400
verts = (rs.standard_normal((B, N, V, 3)) * 0.1 + rs.standard_normal((B, 1, V, 3)) * 0.3).astype(np.float32)
j3d = (rs.standard_normal((B, N, 17, 3)) * 0.05 + rs.standard_normal((B, 1, 17, 3)) * 0.3).astype(np.float32)
j2d_samples = (rs.uniform(40, 216, (B, 1, 17, 2)) + rs.standard_normal((B, N, 17, 2)) * 4).astype(np.float32)
tgt_j2d = rs.uniform(40, 216, (B, 17, 2)).astype(np.float32)
tgt_vis = (rs.uniform(size=(B, 17)) > 0.3)
in_j2d = np.round(rs.uniform(30, 220, (B, 17, 2))).astype(np.float32)
in_vis = (rs.uniform(size=(B, 17)) > 0.4)
heat = convert_2Djoints_to_gaussian_heatmaps_torch(torch.tensor(in_j2d), 256, std=4.0) * torch.tensor(in_vis.astype(np.float32))[:, :, None, None]
model_input = torch.cat([torch.zeros(B, 1, 256, 256), heat], dim=1)
rec_j2d, rec_vis = convert_heatmaps_to_2Djoints_coordinates_torch(joints2D_heatmaps=model_input[:, 1:], eps=1e-6, gaussian_heatmaps=True)
metrics = ['verts3D_sample_diversity', 'joints3D_sample_diversity', 'joints3D_invis_sample_diversity', 'joints3D_vis_sample_diversity',
           'joints2Dsamples-L2E', 'input_joints2Dsamples-L2E']
tr = EvalMetricsTracker(metrics, num_samples_for_prob_metrics=N)
tr.initialise_metric_sums()
tr.initialise_per_frame_metric_lists()
tr.update_per_batch({'verts3D_samples': verts, 'joints3D_coco_samples': j3d, 'joints2Dsamples': j2d_samples},
                    {'joints2D': tgt_j2d, 'joints2D_vis': tgt_vis}, B, model_input=model_input, return_per_frame_metrics=True)
out = {m.replace('-', '_'): np.asarray(tr.per_frame_metrics[m][0], np.float64) for m in metrics}
np.savez(os.path.join(HERE, 'tracker_golden.npz'), verts=verts, j3d=j3d, j2d_samples=j2d_samples, tgt_j2d=tgt_j2d, tgt_vis=tgt_vis,
         in_j2d=rec_j2d.numpy(), in_vis=rec_vis.numpy(), **out)
print({k: v for k, v in out.items()})
