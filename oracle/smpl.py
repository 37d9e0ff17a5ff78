"""Oracle (test infrastructure): SMPL linear blend skinning + the reference's 90-joint assembly.

Reference-owned wrapper followed literally: /root/reference/models/smpl.py:13-41.
[upstream, PARITY UNPINNED] everything inside ``lbs`` / ``SMPL.forward`` restates smplx==0.1.26
(requirements.txt:10; lbs.py ``lbs``, ``blend_shapes``, ``vertices2joints``, ``batch_rigid_transform``,
``transform_mat``; body_models.py ``SMPL.forward``; vertex_joint_selector.py; vertex_ids.py 'smplh').
Model data are passed in as a dict of fp32 tensors:
  v_template (V,3) shapedirs (V,3,10) posedirs (207,3V) J_regressor (24,V) lbs_weights (V,24)
  parents list(24)  J_regressor_extra (9,V) J_regressor_cocoplus (19,V) J_regressor_h36m (17,V)
"""
import torch
import torch.nn.functional as F

from .so3 import batch_rodrigues

# [upstream, from memory] smplx vertex_ids['smplh'] in VertexJointSelector order:
# face (nose, reye, leye, rear, lear), feet (LBigToe, LSmallToe, LHeel, RBigToe, RSmallToe, RHeel),
# finger tips (l then r: thumb, index, middle, ring, pinky).
EXTRA_VERTEX_JOINTS = [332, 6260, 2800, 4071, 583,
                       3216, 3226, 3387, 6617, 6624, 6787,
                       2746, 2319, 2445, 2556, 2673,
                       6191, 5782, 5905, 6016, 6133]

SMPL_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]


def vertices2joints(J_regressor, vertices):
    """[upstream] lbs.vertices2joints: einsum('bik,ji->bjk')."""
    return torch.einsum('bik,ji->bjk', vertices, J_regressor)


def batch_rigid_transform(rot_mats, joints, parents):
    """[upstream] lbs.batch_rigid_transform: relative joints, 4x4 chain by parent, posed joints =
    translation column, rel transforms = T - [0 | T @ [J;0]]."""
    joints = joints.unsqueeze(-1)
    rel = joints.clone()
    rel[:, 1:] = rel[:, 1:] - joints[:, parents[1:]]
    tm = torch.cat([F.pad(rot_mats.reshape(-1, 3, 3), [0, 0, 0, 1]),
                    F.pad(rel.reshape(-1, 3, 1), [0, 0, 0, 1], value=1.0)], dim=2)
    tm = tm.reshape(-1, joints.shape[1], 4, 4)
    chain = [tm[:, 0]]
    for i in range(1, len(parents)):
        chain.append(torch.matmul(chain[parents[i]], tm[:, i]))
    transforms = torch.stack(chain, dim=1)
    posed = transforms[:, :, :3, 3]
    jh = F.pad(joints, [0, 0, 0, 1])
    rel_tf = transforms - F.pad(torch.matmul(transforms, jh), [3, 0, 0, 0, 0, 0, 0, 0])
    return posed, rel_tf


def lbs(betas, pose, model, pose2rot=True):
    """[upstream] lbs.lbs.  pose: (M,72) axis-angle if pose2rot else (M,24,3,3)."""
    M = max(betas.shape[0], pose.shape[0])
    dtype = betas.dtype
    v_shaped = model['v_template'] + torch.einsum('bl,mkl->bmk', betas, model['shapedirs'])
    J = vertices2joints(model['J_regressor'], v_shaped)
    ident = torch.eye(3, dtype=dtype)
    if pose2rot:
        rot_mats = batch_rodrigues(pose.reshape(-1, 3)).view(M, -1, 3, 3)
    else:
        rot_mats = pose.view(M, -1, 3, 3)
    pose_feature = (rot_mats[:, 1:] - ident).reshape(M, -1)
    v_posed = torch.matmul(pose_feature, model['posedirs']).view(M, -1, 3) + v_shaped
    J_tf, A = batch_rigid_transform(rot_mats, J, model['parents'])
    W = model['lbs_weights'].unsqueeze(0).expand(M, -1, -1)
    T = torch.matmul(W, A.view(M, 24, 16)).view(M, -1, 4, 4)
    v_h = torch.cat([v_posed, torch.ones(M, v_posed.shape[1], 1, dtype=dtype)], dim=2)
    verts = torch.matmul(T, v_h.unsqueeze(-1))[:, :, :3, 0]
    return verts, J_tf


def smpl_forward(model, betas, body_pose, global_orient, pose2rot=True, transl=None):
    """models/smpl.py:27-41 on top of [upstream] smplx SMPL.forward:
    lbs -> 24 joints + 21 vertex-picked joints (=45) -> (+transl) -> + 9 extra + 19 cocoplus + 17 h36m
    regressed from the final vertices -> joints (M,90,3)."""
    full_pose = torch.cat([global_orient, body_pose], dim=1)
    M = max(betas.shape[0], global_orient.shape[0], body_pose.shape[0])
    if betas.shape[0] != M:
        betas = betas.expand(int(M / betas.shape[0]), -1)
    verts, joints = lbs(betas, full_pose, model, pose2rot=pose2rot)
    joints = torch.cat([joints, verts[:, EXTRA_VERTEX_JOINTS]], dim=1)
    if transl is not None:
        joints = joints + transl.unsqueeze(1)
        verts = verts + transl.unsqueeze(1)
    extra = vertices2joints(model['J_regressor_extra'], verts)
    coco = vertices2joints(model['J_regressor_cocoplus'], verts)
    h36m = vertices2joints(model['J_regressor_h36m'], verts)
    return verts, torch.cat([joints, extra, coco, h36m], dim=1)
