"""Golden vectors for the mesh post-processing helpers, from the REAL reference functions under /root/reference
(build container only):  utils/sampling_utils.py::compute_vertex_variance_from_samples,
utils/cam_utils.py::orthographic_project_torch, utils/joints2d_utils.py::undo_keypoint_normalisation.
    python tests/golden/make_golden_sampling.py   ->  tests/golden/sampling_golden.npz"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, '/root/reference')
from utils.sampling_utils import compute_vertex_variance_from_samples  # noqa: E402
from utils.cam_utils import orthographic_project_torch  # noqa: E402
from utils.joints2d_utils import undo_keypoint_normalisation  # noqa: E402
from utils.label_conversions import ALL_JOINTS_TO_COCO_MAP  # noqa: E402

g = torch.Generator().manual_seed(0)
verts = torch.randn(25, 300, 3, generator=g) * 0.05 + torch.randn(1, 300, 3, generator=g)
avg, std = compute_vertex_variance_from_samples(verts)
joints = torch.randn(12, 90, 3, generator=g)
cam = torch.cat([torch.rand(3, 1, generator=g) + 0.5, torch.randn(3, 2, generator=g) * 0.1], 1)
sel = joints[:, ALL_JOINTS_TO_COCO_MAP, :]
proj = orthographic_project_torch(sel, cam.repeat_interleave(4, dim=0))
pix = undo_keypoint_normalisation(proj, 256)
np.savez(os.path.join(HERE, 'sampling_golden.npz'), verts=verts.numpy(), avg=avg.numpy(), std=std.numpy(), joints=joints.numpy(),
         cam=cam.numpy(), proj=proj.numpy(), pix=pix.numpy(), coco=np.asarray(ALL_JOINTS_TO_COCO_MAP))
print('wrote sampling_golden.npz', avg.shape, std.shape, proj.shape)
