"""SMPL body model with the reference's 90-joint output, running on the fused sm_100a LBS kernels.

Drop-in for /root/reference/models/smpl.py:13-41 (``SMPL(_SMPL)``) and the parts of smplx 0.1.26
``SMPL.__init__/forward`` that the reference's callers rely on:

    smpl = SMPL(model_path, batch_size=1, gender='neutral', num_betas=10, create_transl=True)
    out = smpl(betas=..., body_pose=..., global_orient=..., pose2rot=False)
    out.vertices (M,6890,3)   out.joints (M,90,3)   out.betas / body_pose / global_orient / full_pose

Compute goes through the C-ABI (``hf_smpl_create`` / ``hf_lbs_forward`` / ``hf_rodrigues``); there is no
CPU path.  Model data: ``<model_path>/SMPL_<GENDER>.pkl`` as smplx expects (licence-gated, not shipped),
or ``SMPL.from_arrays`` for a dict of arrays (used with ``humaniflow_b200.synthetic``).
"""
import ctypes
import os
import pickle
from collections import namedtuple

import numpy as np
import torch
from torch import nn

from . import _lib

SMPLOutput = namedtuple('SMPLOutput', ['vertices', 'joints', 'full_pose', 'betas', 'global_orient', 'body_pose'])

# [upstream, from memory] smplx vertex_ids['smplh'] in VertexJointSelector order (face, feet, finger tips l/r)
VERTEX_JOINT_IDS = [332, 6260, 2800, 4071, 583, 3216, 3226, 3387, 6617, 6624, 6787,
                    2746, 2319, 2445, 2556, 2673, 6191, 5782, 5905, 6016, 6133]

# reference configs/paths.py:3-5
J_REGRESSOR_EXTRA = './model_files/J_regressor_extra.npy'
COCOPLUS_REGRESSOR = './model_files/cocoplus_regressor.npy'
H36M_REGRESSOR = './model_files/J_regressor_h36m.npy'


def _to_np(a):
    if hasattr(a, 'todense'):
        a = a.todense()
    if hasattr(a, 'r'):          # chumpy array
        a = a.r
    return np.asarray(a)


def load_smpl_pkl(model_path, gender):
    """Read an SMPL .pkl the way smplx does ([upstream] body_models.SMPL.__init__)."""
    if os.path.isdir(model_path):
        model_path = os.path.join(model_path, 'SMPL_%s.pkl' % gender.upper())
    if not os.path.exists(model_path):
        raise FileNotFoundError('SMPL model file %s not found (licence-gated download, see the reference README); '
                                'use SMPL.from_arrays(...) for synthetic model data' % model_path)
    with open(model_path, 'rb') as f:
        d = pickle.load(f, encoding='latin1')
    V = _to_np(d['v_template']).shape[0]
    return {
        'v_template': _to_np(d['v_template']),
        'shapedirs': _to_np(d['shapedirs']),
        'posedirs': _to_np(d['posedirs']).reshape(V * 3, -1).T,
        'J_regressor': _to_np(d['J_regressor']),
        'lbs_weights': _to_np(d['weights']),
        'parents': [int(p) for p in np.asarray(d['kintree_table'])[0].astype(np.int64)],
        'faces': _to_np(d['f']).astype(np.int64) if 'f' in d else None,
    }


class _LBSFunction(torch.autograd.Function):
    """Differentiable LBS: forward = hf_lbs_forward, backward = hf_lbs_backward.  The reference gets the
    gradient from torch.autograd through smplx's lbs (models/smpl.py:27-41); here it is three launches of lbs.cu."""

    @staticmethod
    def forward(ctx, smpl, betas, rotmats, transl):
        verts, joints = smpl._lbs_raw(betas, rotmats, transl)
        ctx.set_materialize_grads(False)          # an unused output arrives as None, not as 80 KB of zeros per body
        ctx.smpl = smpl
        ctx.save_for_backward(betas, rotmats)
        ctx.has_transl = transl is not None
        return verts, joints

    @staticmethod
    def backward(ctx, g_verts, g_joints):
        betas, rotmats = ctx.saved_tensors
        smpl = ctx.smpl
        if g_verts is None and g_joints is None:
            return None, None, None, None
        lib = _lib.load()
        dev = betas.device
        M = rotmats.shape[0]
        gv = None if g_verts is None else _lib.f32c(g_verts)
        gj = None if g_joints is None else _lib.f32c(g_joints)
        g_betas = torch.empty_like(betas)
        g_rot = torch.empty_like(rotmats)
        h = smpl._handle(dev)
        with torch.cuda.device(dev):
            nbytes = lib.hf_lbs_backward_workspace_bytes(h, M)
            ws = smpl._ws_bwd.get(dev)
            if ws is None or ws.numel() < nbytes:
                ws = torch.empty(max(nbytes, 1), device=dev, dtype=torch.uint8)
                smpl._ws_bwd[dev] = ws
            _lib.check(lib.hf_lbs_backward(h, _lib.ptr(betas), _lib.ptr(rotmats), _lib.ptr(gv), _lib.ptr(gj), _lib.ptr(g_betas),
                                           _lib.ptr(g_rot), _lib.ptr(ws), ws.numel(), M, _lib.stream()))
        g_transl = None
        if ctx.has_transl and ctx.needs_input_grad[3]:
            g_transl = torch.zeros(M, 3, device=dev)
            if gv is not None:
                g_transl = g_transl + gv.sum(1)
            if gj is not None:
                g_transl = g_transl + gj.sum(1)
        return None, g_betas, g_rot, g_transl


def _rodrigues_torch(aa):
    """(n,3) axis-angle -> (n,3,3), differentiable (torch ops; [upstream] smplx lbs.batch_rodrigues, same 1e-8 guard).  Only used
    when a gradient is required through pose2rot=True; the inference path uses hf_rodrigues."""
    angle = torch.norm(aa + 1e-8, dim=1, keepdim=True)
    d = aa / angle
    c, s_ = torch.cos(angle)[:, None], torch.sin(angle)[:, None]
    rx, ry, rz = d[:, 0:1], d[:, 1:2], d[:, 2:3]
    z = torch.zeros_like(rx)
    K = torch.cat([z, -rz, ry, rz, z, -rx, -ry, rx, z], dim=1).view(-1, 3, 3)
    eye = torch.eye(3, dtype=aa.dtype, device=aa.device)[None]
    return eye + s_ * K + (1 - c) * torch.bmm(K, K)


class SMPL(nn.Module):
    NUM_JOINTS = 23
    NUM_BODY_JOINTS = 23

    def __init__(self, model_path=None, batch_size=1, gender='neutral', num_betas=10, create_transl=True,
                 create_betas=True, create_global_orient=True, create_body_pose=True, data=None,
                 regressors=None, vertex_joint_ids=None, **kwargs):
        super().__init__()
        if data is None:
            data = load_smpl_pkl(model_path, gender)
        self.gender = gender
        self.batch_size = batch_size
        self.num_betas = num_betas
        f32 = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float32) if not torch.is_tensor(a) else a.float())
        shapedirs = f32(data['shapedirs'])[:, :, :num_betas]
        parents = [int(p) for p in data['parents']]
        parents[0] = -1
        self.register_buffer('v_template', f32(data['v_template']))
        self.register_buffer('shapedirs', shapedirs.contiguous())
        self.register_buffer('posedirs', f32(data['posedirs']))
        self.register_buffer('J_regressor', f32(data['J_regressor']))
        self.register_buffer('lbs_weights', f32(data['lbs_weights']))
        self.register_buffer('parents', torch.tensor(parents, dtype=torch.long))
        if data.get('faces') is not None:
            self.register_buffer('faces_tensor', torch.as_tensor(np.asarray(data['faces'], dtype=np.int64)))
            self.faces = np.asarray(data['faces'])
        # models/smpl.py:16-25: the three extra regressors
        if regressors is None:
            regressors = {}
            for key, path in (('J_regressor_extra', J_REGRESSOR_EXTRA), ('J_regressor_cocoplus', COCOPLUS_REGRESSOR),
                              ('J_regressor_h36m', H36M_REGRESSOR)):
                regressors[key] = data[key] if key in data else np.load(path)
        for key in ('J_regressor_extra', 'J_regressor_cocoplus', 'J_regressor_h36m'):
            self.register_buffer(key, f32(regressors[key]))
        self.vertex_joint_ids = list(VERTEX_JOINT_IDS if vertex_joint_ids is None else vertex_joint_ids)
        # [upstream] default parameters of size batch_size
        if create_betas:
            self.betas = nn.Parameter(torch.zeros(batch_size, num_betas))
        if create_global_orient:
            self.global_orient = nn.Parameter(torch.zeros(batch_size, 3))
        if create_body_pose:
            self.body_pose = nn.Parameter(torch.zeros(batch_size, self.NUM_BODY_JOINTS * 3))
        if create_transl:
            self.transl = nn.Parameter(torch.zeros(batch_size, 3))
        self._handles = {}
        self._ws = {}
        self._ws_bwd = {}

    @classmethod
    def from_arrays(cls, data, batch_size=1, num_betas=10, create_transl=True, **kwargs):
        regs = {k: data[k] for k in ('J_regressor_extra', 'J_regressor_cocoplus', 'J_regressor_h36m')}
        return cls(data=data, regressors=regs, batch_size=batch_size, num_betas=num_betas,
                   create_transl=create_transl, **kwargs)

    @property
    def num_joints_out(self):
        return self.J_regressor.shape[0] + len(self.vertex_joint_ids) + self.J_regressor_extra.shape[0] + \
            self.J_regressor_cocoplus.shape[0] + self.J_regressor_h36m.shape[0]

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._handles = {}
        return out

    def _handle(self, device):
        """Register the (immutable) model data with the CUDA library for this device."""
        key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
        if key in self._handles:
            return self._handles[key]
        lib = _lib.load()
        cpu = lambda t: t.detach().to('cpu', torch.float32).contiguous()
        vt, sd, pd, jr, lw = cpu(self.v_template), cpu(self.shapedirs), cpu(self.posedirs), cpu(self.J_regressor), cpu(self.lbs_weights)
        extra = torch.cat([cpu(self.J_regressor_extra), cpu(self.J_regressor_cocoplus), cpu(self.J_regressor_h36m)], 0).contiguous()
        parents = self.parents.detach().cpu().to(torch.int32).contiguous()
        vj = torch.tensor(self.vertex_joint_ids, dtype=torch.int32)
        h = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.hf_smpl_create(ctypes.byref(h), vt.shape[0], sd.shape[2], jr.shape[0], _lib.ptr(vt), _lib.ptr(sd),
                                          _lib.ptr(pd), _lib.ptr(jr), _lib.ptr(lw), _lib.ptr(parents), _lib.ptr(vj),
                                          len(self.vertex_joint_ids), _lib.ptr(extra), extra.shape[0]))
        if getattr(self, '_impl', 0):
            _lib.check(lib.hf_lbs_set_impl(h, self._impl))
        self._handles[key] = h
        return h

    def set_impl(self, impl):
        """0 = persistent fp16 tcgen05 blend (product path), 1 = FP32 CUDA-core blend, 2 = split-bf16 three-pass tcgen05
        blend (both cross-checks), 3 = lanes-as-samples fp16 blend (round-2 experiment)."""
        self._impl = impl
        for h in self._handles.values():
            _lib.check(_lib.load().hf_lbs_set_impl(h, impl))

    def lbs(self, betas, rotmats, transl=None, out_vertices=None, out_joints=None):
        """betas (M,nb), rotmats (M,24,3,3) fp32 CUDA -> vertices (M,V,3), joints (M,90,3).
        ``out_vertices`` / ``out_joints``: optional preallocated fp32 CUDA outputs (serving loops that double-buffer).
        When autograd is recording and an input requires a gradient, the call goes through _LBSFunction (CUDA backward)."""
        if out_vertices is None and out_joints is None and torch.is_grad_enabled() and \
                any(torch.is_tensor(t) and t.requires_grad for t in (betas, rotmats, transl)):
            _lib.require_cuda('SMPL.forward')
            if not rotmats.is_cuda:
                raise RuntimeError('humaniflow_b200.SMPL: inputs must be CUDA tensors (no CPU fallback)')
            M = rotmats.shape[0]
            keep = lambda t: t.to(torch.float32).contiguous()          # stays on the autograd tape (f32c detaches)
            betas = keep(betas)
            if betas.shape[0] != M:
                betas = betas.expand(int(M / betas.shape[0]), -1).contiguous()
            rotmats = keep(rotmats)
            transl = None if transl is None else keep(transl).expand(M, 3).contiguous()
            return _LBSFunction.apply(self, betas, rotmats, transl)
        return self._lbs_raw(betas, rotmats, transl, out_vertices, out_joints)

    def _lbs_raw(self, betas, rotmats, transl=None, out_vertices=None, out_joints=None):
        _lib.require_cuda('SMPL.forward')
        if not betas.is_cuda:
            raise RuntimeError('humaniflow_b200.SMPL: inputs must be CUDA tensors (no CPU fallback)')
        lib = _lib.load()
        dev = betas.device
        M = rotmats.shape[0]
        betas = _lib.f32c(betas)
        if betas.shape[0] != M:
            betas = betas.expand(int(M / betas.shape[0]), -1).contiguous()     # [upstream] SMPL.forward
        rotmats = _lib.f32c(rotmats)
        transl = None if transl is None else _lib.f32c(transl).expand(M, 3).contiguous()
        h = self._handle(dev)
        V = self.v_template.shape[0]
        verts = torch.empty(M, V, 3, device=dev, dtype=torch.float32) if out_vertices is None else out_vertices
        joints = torch.empty(M, self.num_joints_out, 3, device=dev, dtype=torch.float32) if out_joints is None else out_joints
        assert verts.shape == (M, V, 3) and verts.is_contiguous() and verts.dtype == torch.float32 and verts.device == dev
        assert joints.shape == (M, self.num_joints_out, 3) and joints.is_contiguous() and joints.dtype == torch.float32
        with torch.cuda.device(dev):
            nbytes = lib.hf_lbs_workspace_bytes(h, M)
            ws = self._ws.get(dev)
            if ws is None or ws.numel() < nbytes:
                ws = torch.empty(max(nbytes, 1), device=dev, dtype=torch.uint8)
                self._ws[dev] = ws
            _lib.check(lib.hf_lbs_forward(h, _lib.ptr(betas), _lib.ptr(rotmats), _lib.ptr(transl), _lib.ptr(verts),
                                          _lib.ptr(joints), _lib.ptr(ws), ws.numel(), M, _lib.stream()))
        return verts, joints

    def forward_samples(self, betas, body_rotmats, glob_rotmats, samples_per_image, transl=None, out_vertices=None, out_joints=None):
        """The sampled bodies of a batch in one call, with the rotations as HumaniflowModel returns them: betas (B*N,nb),
        body_rotmats (B*N,23,3,3), glob_rotmats (B,3,3) shared by the N samples of each image (predict_humaniflow.py:138-141
        expands and concatenates them first).  Returns an SMPLOutput with vertices (B*N,V,3) and joints (B*N,90,3)."""
        _lib.require_cuda('SMPL.forward_samples')
        if not betas.is_cuda:
            raise RuntimeError('humaniflow_b200.SMPL: inputs must be CUDA tensors (no CPU fallback)')
        lib = _lib.load()
        dev = betas.device
        betas, body, glob = _lib.f32c(betas), _lib.f32c(body_rotmats), _lib.f32c(glob_rotmats)
        M, N = body.shape[0], int(samples_per_image)
        assert betas.shape[0] == M and glob.shape[0] * N == M and body.shape[1] == self.NUM_BODY_JOINTS
        transl = None if transl is None else _lib.f32c(transl).expand(M, 3).contiguous()
        h = self._handle(dev)
        V = self.v_template.shape[0]
        verts = torch.empty(M, V, 3, device=dev, dtype=torch.float32) if out_vertices is None else out_vertices
        joints = torch.empty(M, self.num_joints_out, 3, device=dev, dtype=torch.float32) if out_joints is None else out_joints
        with torch.cuda.device(dev):
            nbytes = lib.hf_lbs_workspace_bytes(h, M)
            ws = self._ws.get(dev)
            if ws is None or ws.numel() < nbytes:
                ws = torch.empty(max(nbytes, 1), device=dev, dtype=torch.uint8)
                self._ws[dev] = ws
            _lib.check(lib.hf_lbs_forward_split(h, _lib.ptr(betas), _lib.ptr(body), _lib.ptr(glob), N, _lib.ptr(transl), _lib.ptr(verts),
                                                _lib.ptr(joints), _lib.ptr(ws), ws.numel(), M, _lib.stream()))
        return SMPLOutput(vertices=verts, joints=joints, full_pose=None, betas=betas, global_orient=glob, body_pose=body)

    def tpose(self, betas=None, transl=None):
        """T-pose meshes: what ``forward(betas=...)`` returns with the default zero pose (predict_humaniflow.py:147,
        evaluate_humaniflow.py:131-133), computed without pose blend and skinning.  betas (M,nb) CUDA fp32."""
        _lib.require_cuda('SMPL.tpose')
        betas = betas if betas is not None else self.betas
        if not betas.is_cuda:
            raise RuntimeError('humaniflow_b200.SMPL: inputs must be CUDA tensors (no CPU fallback)')
        lib = _lib.load()
        betas = _lib.f32c(betas)
        M, dev = betas.shape[0], betas.device
        transl = None if transl is None else _lib.f32c(transl).expand(M, 3).contiguous()
        h = self._handle(dev)
        verts = torch.empty(M, self.v_template.shape[0], 3, device=dev, dtype=torch.float32)
        joints = torch.empty(M, self.num_joints_out, 3, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(lib.hf_lbs_tpose(h, _lib.ptr(betas), _lib.ptr(transl), _lib.ptr(verts), _lib.ptr(joints), M, _lib.stream()))
        return SMPLOutput(vertices=verts, joints=joints, full_pose=None, betas=betas, global_orient=None, body_pose=None)

    def forward(self, betas=None, body_pose=None, global_orient=None, transl=None, return_verts=True,
                return_full_pose=False, pose2rot=True, out_vertices=None, out_joints=None, **kwargs):
        """models/smpl.py:27-41 + [upstream] smplx SMPL.forward.  ``None`` inputs fall back to the module's
        parameters; ``pose2rot=False`` takes rotation matrices (M,23,3,3)/(M,1,3,3), ``True`` axis-angle."""
        global_orient = global_orient if global_orient is not None else self.global_orient
        body_pose = body_pose if body_pose is not None else self.body_pose
        betas = betas if betas is not None else self.betas
        if transl is None and hasattr(self, 'transl'):
            transl = self.transl
        M = max(betas.shape[0], global_orient.shape[0], body_pose.shape[0])
        lib = _lib.load()
        if pose2rot:
            _lib.require_cuda('SMPL.forward')
            full_pose = torch.cat([global_orient.reshape(global_orient.shape[0], -1),
                                   body_pose.reshape(body_pose.shape[0], -1)], dim=1)
            if not full_pose.is_cuda:
                raise RuntimeError('humaniflow_b200.SMPL: inputs must be CUDA tensors (no CPU fallback)')
            want_grad = torch.is_grad_enabled() and full_pose.requires_grad and out_vertices is None and out_joints is None
            aa = (full_pose.to(torch.float32) if want_grad else _lib.f32c(full_pose)).reshape(-1, 3)
            if want_grad:
                rot = _rodrigues_torch(aa)             # a gradient w.r.t. the axis-angle pose is wanted (fitting loops)
            else:
                rot = torch.empty(aa.shape[0], 3, 3, device=aa.device, dtype=torch.float32)
                with torch.cuda.device(aa.device):
                    _lib.check(lib.hf_rodrigues(_lib.ptr(aa), _lib.ptr(rot), aa.shape[0], _lib.stream()))
            rotmats = rot.view(full_pose.shape[0], -1, 3, 3)
        else:
            full_pose = torch.cat([global_orient.reshape(-1, 1, 3, 3), body_pose.reshape(body_pose.shape[0], -1, 3, 3)], dim=1)
            rotmats = full_pose
        if rotmats.shape[0] != M:
            rotmats = rotmats.expand(M, -1, -1, -1)
        if transl is not None and transl.shape[0] not in (1, M):
            raise ValueError('transl batch %d does not match %d' % (transl.shape[0], M))
        verts, joints = self.lbs(betas.to(rotmats.device), rotmats, transl, out_vertices, out_joints)
        return SMPLOutput(vertices=verts if return_verts else None, joints=joints,
                          full_pose=full_pose if return_full_pose else None, betas=betas,
                          global_orient=global_orient, body_pose=body_pose)

    def __del__(self):
        try:
            lib = _lib.load()
            for h in self._handles.values():
                lib.hf_smpl_destroy(h)
        except Exception:
            pass
