"""Oracle (test infrastructure): CPU restatement of the reference helpers that post-process the sampled meshes.

PINNED by tests/golden/sampling_golden.npz, produced by the real reference functions
(utils/sampling_utils.py, utils/cam_utils.py, utils/joints2d_utils.py) in the build container.
The 180-degree flip goes through pytorch3d in the reference (utils/rigid_transform_utils.py:67-83,
``so3_exp_pytorch3d``), which is not installable here: restated as the exact Rodrigues rotation by pi about x,
diag(1,-1,-1) [upstream, unpinned; pytorch3d's fp32 sin(pi) differs by ~9e-8].
"""
import torch

ALL_JOINTS_TO_COCO_MAP = [24, 26, 25, 28, 27, 16, 17, 18, 19, 20, 21, 1, 2, 4, 5, 7, 8]   # utils/label_conversions.py:17


def compute_vertex_variance_from_samples(vertices_samples):
    """utils/sampling_utils.py:22-33.  (N,V,3) samples of one image -> (mean distance of a vertex from its sample mean (V,),
    per-coordinate RMS deviation (V,3))."""
    dev = vertices_samples - vertices_samples.mean(dim=0)
    rms_per_coordinate = (dev ** 2).mean(dim=0).sqrt()
    mean_distance = dev.norm(dim=-1).mean(dim=0)
    return mean_distance, rms_per_coordinate


def orthographic_project(points3D, cam_params):
    """utils/cam_utils.py:9-16.  (B,N,3), (B,3) -> (B,N,2)."""
    return cam_params[:, None, [0]] * (points3D[:, :, :2] + cam_params[:, None, 1:])


def undo_keypoint_normalisation(normalised_keypoints, img_wh):
    """utils/joints2d_utils.py:5-10."""
    return (normalised_keypoints + 1) * (img_wh / 2.0)


def project_joints2d(joints, cam_wp, joint_ids=ALL_JOINTS_TO_COCO_MAP, flip_x=True, img_wh=None):
    """utils/sampling_utils.py:50-58: joints (M,J,3) of M = B*n samples, cam_wp (B,3)."""
    j = joints[:, list(joint_ids), :] if joint_ids is not None else joints
    if flip_x:
        j = j * torch.tensor([1.0, -1.0, -1.0], dtype=j.dtype)
    per = j.shape[0] // cam_wp.shape[0]
    cam = cam_wp.repeat_interleave(per, dim=0)
    out = orthographic_project(j, cam)
    return undo_keypoint_normalisation(out, img_wh) if img_wh else out
