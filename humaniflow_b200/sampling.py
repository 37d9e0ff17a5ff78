"""Per-image reductions / projections of the sampled meshes: CUDA mirrors of the reference's
``utils/sampling_utils.py`` helpers that run right after the SMPL forward in predict / evaluate.

    compute_vertex_variance_from_samples   utils/sampling_utils.py:22-33 (predict_humaniflow.py:168)
    project_joints2d                       utils/sampling_utils.py:50-58, evaluate_humaniflow.py:186-206
                                           (= aa_rotate_translate_points_pytorch3d(x, pi) + orthographic_project_torch
                                              + undo_keypoint_normalisation on selected joints)

No CPU fallback: CUDA tensors only.
"""
import torch

from . import _lib

ALL_JOINTS_TO_COCO_MAP = [24, 26, 25, 28, 27, 16, 17, 18, 19, 20, 21, 1, 2, 4, 5, 7, 8]   # utils/label_conversions.py:17


def compute_vertex_variance_from_samples(vertices_samples):
    """vertices_samples (N,V,3) -> (avg_vertex_l2_distance_from_mean (V,), directional_vertex_variances (V,3)), the
    reference's signature; a batched (B,N,V,3) input returns (B,V) and (B,V,3)."""
    _lib.require_cuda('compute_vertex_variance_from_samples')
    if not vertices_samples.is_cuda:
        raise RuntimeError('humaniflow_b200.sampling: inputs must be CUDA tensors (no CPU fallback)')
    x = _lib.f32c(vertices_samples)
    batched = x.dim() == 4
    if not batched:
        x = x[None]
    B, N, V, _ = x.shape
    avg = torch.empty(B, V, device=x.device, dtype=torch.float32)
    std = torch.empty(B, V, 3, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().hf_vertex_variance(_lib.ptr(x), B, N, V, _lib.ptr(avg), _lib.ptr(std), _lib.stream()))
    return (avg, std) if batched else (avg[0], std[0])


def project_joints2d(joints, cam_wp, joint_ids=ALL_JOINTS_TO_COCO_MAP, flip_x=True, img_wh=None):
    """joints (M,J,3), cam_wp (B,3) with M = B * samples_per_image -> (M,len(joint_ids),2): the selected joints, rotated
    by pi about x (``flip_x``), weak-perspective projected with their image's camera and, if ``img_wh`` is given,
    mapped from [-1,1] to pixels."""
    _lib.require_cuda('project_joints2d')
    if not joints.is_cuda:
        raise RuntimeError('humaniflow_b200.sampling: inputs must be CUDA tensors (no CPU fallback)')
    j = _lib.f32c(joints)
    cam = _lib.f32c(cam_wp).to(j.device)
    M, J = j.shape[0], j.shape[1]
    if M % cam.shape[0]:
        raise ValueError('%d joint sets do not divide over %d cameras' % (M, cam.shape[0]))
    ids = None if joint_ids is None else torch.tensor(list(joint_ids), dtype=torch.int32, device=j.device)
    n = J if ids is None else ids.numel()
    out = torch.empty(M, n, 2, device=j.device, dtype=torch.float32)
    with torch.cuda.device(j.device):
        _lib.check(_lib.load().hf_project_joints2d(_lib.ptr(j), _lib.ptr(cam), _lib.ptr(ids), M, M // cam.shape[0], J, n,
                                                   int(bool(flip_x)), float(img_wh or 0.0), _lib.ptr(out), _lib.stream()))
    return out
