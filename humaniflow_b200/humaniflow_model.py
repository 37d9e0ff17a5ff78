"""HumaniflowModel on hand-written sm_100a kernels, behind the reference's interface.

Drop-in for /root/reference/models/humaniflow_model.py:33-340: same constructor
``HumaniflowModel(device, model_cfg, smpl_parents)``, same ``forward`` keyword arguments, same return-dict
keys/shapes/dtypes, same state-dict key names (``image_encoder.*``, ``fc1``, ``fc_shape``, ``fc_glob``,
``fc_cam``, ``fc_input_shape_glob_cam_feats``, ``fc_flow_context.{j}``,
``pose_so3flow_transform_modules.{2j+t}.nn.layers.{l}``, buffers ``init_glob`` / ``init_cam``), so a
reference checkpoint loads with ``strict=True``.

What runs where: torch holds parameters and allocates outputs; every arithmetic step of the path is a
kernel of libhumaniflow_b200.so called through the C-ABI (include/humaniflow_b200.h).  Inference only:
outputs carry no autograd graph (the backward of this path is SURVEY.md 8f N3).  Two keyword-only
extensions make results reproducible against the CPU oracle: ``base_noise`` (B,N,23,3), the
base-distribution draws, and ``shape_eps`` (B,N,10); when omitted they are drawn from torch's generator.
"""
import ctypes
from collections import defaultdict

import torch
from torch import nn
from torch.distributions import Normal

from . import _lib
from .resnet import resnet18, resnet50


def immediate_parent_to_all_ancestors(immediate_parents):
    """models/humaniflow_model.py:16-30: ancestors of every non-root joint, nearest first, root excluded
    (joint 0 here is SMPL joint 1)."""
    ancestors = defaultdict(list)
    for i in range(1, len(immediate_parents)):
        joint, parent = i - 1, immediate_parents[i] - 1
        if parent >= 0:
            ancestors[joint] += [parent] + ancestors[parent]
    return ancestors


class _DenseNN(nn.Module):
    """Parameter container with pyro ConditionalDenseNN's layer naming ([upstream] ``.layers.{l}``)."""

    def __init__(self, input_dim, context_dim, hidden_dims, out_dim):
        super().__init__()
        dims = [input_dim + context_dim] + list(hidden_dims) + [out_dim]
        self.layers = nn.ModuleList([nn.Linear(dims[i], dims[i + 1]) for i in range(len(dims) - 1)])


class ConditionalSplineCoupling(nn.Module):
    """Parameter container for one coupling (transforms/conditional_spline_coupling_transform.py:10-78);
    the arithmetic lives in the fused flow kernels."""

    def __init__(self, input_dim, context_dim, hidden_dims, count_bins, bound):
        super().__init__()
        split = input_dim // 2
        n = input_dim - split
        self.nn = _DenseNN(split, context_dim, hidden_dims, n * count_bins * 3 + n * (count_bins - 1))
        self.count_bins, self.bound = count_bins, bound


class _FlowLogProbFunction(torch.autograd.Function):
    """log_prob of joints [joint_first, joint_first + joint_count) with a CUDA backward w.r.t. the contexts and the targets
    (hf_flow_log_prob_backward / hf_flow_algebra_log_prob_backward; weights are constants here)."""

    @staticmethod
    def forward(ctx_, model, joint_first, joint_count, on_group, ctx_all, value):
        lib = _lib.load()
        R = ctx_all.shape[0]
        dev = ctx_all.device
        c = _lib.f32c(ctx_all)
        v = value.detach().to(dev, torch.float64 if on_group else torch.float32).contiguous()
        out = torch.empty(R, joint_count, device=dev, dtype=torch.float32)
        stride = c.shape[1] * c.shape[2]
        with torch.cuda.device(dev):
            fn = lib.hf_flow_log_prob if on_group else lib.hf_flow_algebra_log_prob
            _lib.check(fn(model._flow, _lib.ptr(c), stride, joint_first, joint_count, _lib.ptr(v), R, _lib.ptr(out), _lib.stream()))
        ctx_.model, ctx_.jf, ctx_.jc, ctx_.on_group, ctx_.vdtype = model, joint_first, joint_count, on_group, value.dtype
        ctx_.save_for_backward(c, v)
        return out

    @staticmethod
    def backward(ctx_, g):
        lib = _lib.load()
        c, v = ctx_.saved_tensors
        model, jf, jc = ctx_.model, ctx_.jf, ctx_.jc
        R, dev = c.shape[0], c.device
        g = _lib.f32c(g)
        g_part = torch.empty(R, jc, c.shape[2], device=dev, dtype=torch.float32)
        g_val = torch.empty_like(v)
        stride = c.shape[1] * c.shape[2]
        with torch.cuda.device(dev):
            fn = lib.hf_flow_log_prob_backward if ctx_.on_group else lib.hf_flow_algebra_log_prob_backward
            _lib.check(fn(model._flow, _lib.ptr(c), stride, jf, jc, _lib.ptr(v), _lib.ptr(g), R, _lib.ptr(g_part), _lib.ptr(g_val), _lib.stream()))
        g_ctx = torch.zeros_like(c)
        g_ctx[:, jf:jf + jc] = g_part
        return None, None, None, None, g_ctx, g_val.to(ctx_.vdtype)


class _FlowContextFunction(torch.autograd.Function):
    """Teacher-forced contexts (models/humaniflow_model.py:133-148, 277-283) with a backward w.r.t. the image features, the camera,
    the shape, the ancestors' rotations and the global rotation.  Forward: the library's kernels.  Backward: a handful of small
    dense products with the model's own weight tensors (torch matmuls: glue, 32 x 256-sized)."""

    @staticmethod
    def forward(ctx_, model, input_feats, cam, shape, pose_R, glob_R):
        lib = _lib.load()
        dev = input_feats.device
        B, J = input_feats.shape[0], model.num_bodyparts
        base_ll = model._img_base(input_feats, glob_R, cam)
        betas_ll = _lib.f32c(shape, dev)
        anc_R = _lib.f32c(pose_R, dev)
        assert anc_R.shape == (B, J, 3, 3)
        out = torch.empty(B, J, model.cfg.NORM_FLOW.CONTEXT_DIM, device=dev, dtype=torch.float32)
        idx = model._img_index(B, 1, False, dev)
        with torch.cuda.device(dev):
            _lib.check(lib.hf_flow_context(model._flow, _lib.ptr(base_ll), _lib.ptr(betas_ll), _lib.ptr(idx), _lib.ptr(anc_R),
                                           B, _lib.ptr(out), _lib.stream()))
        ctx_.model = model
        ctx_.dtypes = (input_feats.dtype, cam.dtype, shape.dtype, pose_R.dtype, glob_R.dtype)
        ctx_.save_for_backward(base_ll, betas_ll, out)
        return out

    @staticmethod
    def backward(ctx_, g_ctx):
        model = ctx_.model
        base_ll, betas, out = ctx_.saved_tensors
        B, J = out.shape[0], out.shape[1]
        Fd = model.cfg.INPUT_SHAPE_GLOB_CAM_FEATS_DIM
        Fin, nb = model.input_feats_dim, model.num_shape_params
        with torch.no_grad():
            one = torch.ones((), device=out.device)
            g_pre = g_ctx.float() * torch.where(out > 0, one, out + 1)             # ELU'(pre) = 1 or exp(pre) = ELU(pre) + 1
            g_F = torch.zeros(B, Fd, device=out.device)
            g_pose = torch.zeros(B, J, 3, 3, device=out.device)
            for j in range(J):
                W = model.fc_flow_context[j].weight.detach().float()
                ga = g_pre[:, j] @ W                                                # (B, Fd + 9 * ancestors)
                g_F += ga[:, :Fd]
                for q, a in enumerate(model.ancestors_dict[j]):
                    g_pose[:, a] += ga[:, Fd + 9 * q:Fd + 9 * q + 9].view(B, 3, 3)
            Wimg = model.fc_input_shape_glob_cam_feats.weight.detach().float()
            Wf, Wb, Wg, Wc = Wimg[:, :Fin], Wimg[:, Fin:Fin + nb], Wimg[:, Fin + nb:Fin + nb + 9], Wimg[:, Fin + nb + 9:]
            preF = base_ll + betas @ Wb.t()
            g_preF = g_F * torch.where(preF > 0, one, torch.exp(preF))
            dt = ctx_.dtypes
            return (None, (g_preF @ Wf).to(dt[0]), (g_preF @ Wc).to(dt[1]), (g_preF @ Wb).to(dt[2]), g_pose.to(dt[3]),
                    (g_preF @ Wg).view(B, 3, 3).to(dt[4]))


class ConditionedSO3FlowDist:
    """One joint's flow conditioned on a batch of contexts: the object the reference returns in
    ``conditioned_pose_SO3flow_dists_for_loglik`` (group=True) / ``..._so3flow_...`` (group=False).
    ``log_prob`` follows local_diffeo_transformed_distribution.py:84-142 on the GPU."""

    def __init__(self, model, joint, ctx_all, on_group):
        self.model, self.joint, self.ctx_all, self.on_group = model, joint, ctx_all, on_group
        self._token = model._packed_version          # the packed weights these contexts were computed with

    def log_prob(self, value):
        m, lib = self.model, _lib.load()
        if m._flow is None or m._packed_version is not self._token:
            raise RuntimeError('humaniflow_b200: the model was re-packed (weights changed or moved) after this conditioned '
                               'distribution was created; call forward(compute_for_loglik=True) again')
        R = self.ctx_all.shape[0]
        if torch.is_grad_enabled() and (self.ctx_all.requires_grad or (torch.is_tensor(value) and value.requires_grad)):
            # a gradient w.r.t. the contexts and / or the target is wanted (fitting loops): CUDA forward + CUDA backward
            return _FlowLogProbFunction.apply(m, self.joint, 1, self.on_group, self.ctx_all, value.to(self.ctx_all.device))[:, 0]
        out = torch.empty(R, device=self.ctx_all.device, dtype=torch.float32)
        with torch.cuda.device(out.device):
            if self.on_group:
                v = value.detach().to(self.ctx_all.device, torch.float64).contiguous()
                assert v.shape == (R, 3, 3)
                _lib.check(lib.hf_flow_log_prob(m._flow, _lib.ptr(_lib.f32c(self.ctx_all)), self.ctx_all.shape[1] * self.ctx_all.shape[2],
                                                self.joint, 1, _lib.ptr(v), R, _lib.ptr(out), _lib.stream()))
            else:
                v = _lib.f32c(value, self.ctx_all.device)
                assert v.shape == (R, 3)
                _lib.check(lib.hf_flow_algebra_log_prob(m._flow, _lib.ptr(_lib.f32c(self.ctx_all)), self.ctx_all.shape[1] * self.ctx_all.shape[2],
                                                        self.joint, 1, _lib.ptr(v), R, _lib.ptr(out), _lib.stream()))
        return out


class _FlowDistHandle:
    """Element of ``model.pose_SO3flow_dists``; callers only use ``clear_cache()`` (train_humaniflow.py:353-354)."""

    def clear_cache(self):
        pass


class HumaniflowModel(nn.Module):
    def __init__(self, device, model_cfg, smpl_parents):
        super().__init__()
        self.parents = list(smpl_parents)
        self.ancestors_dict = immediate_parent_to_all_ancestors(self.parents)
        self.num_bodyparts = len(self.parents) - 1
        self.cfg = model_cfg
        self.num_shape_params = model_cfg.NUM_SMPL_BETAS
        self.num_glob_params = 6
        self.register_buffer('init_glob', torch.eye(3)[None, :, :2].contiguous().view(-1, 6).float())   # rotmat_to_rot6d(I)
        self.num_cam_params = 3
        self.register_buffer('init_cam', torch.tensor([0.9, 0.0, 0.0]).float())
        nf0 = model_cfg.NORM_FLOW
        if len(nf0.TRANSFORM_NN_HIDDEN_DIMS) != 3 or nf0.NUM_TRANSFORMS != 2:
            raise NotImplementedError('the flow kernels are specialised to the reference default: 2 transforms per joint, '
                                      '3 hidden layers (configs/humaniflow_config.py:16-19)')
        if model_cfg.NUM_RESNET_LAYERS == 18:
            self.image_encoder = resnet18(in_channels=model_cfg.NUM_IN_CHANNELS, pretrained=False)
            input_feats_dim, fc1_dim = 512, 512
        elif model_cfg.NUM_RESNET_LAYERS == 50:
            self.image_encoder = resnet50(in_channels=model_cfg.NUM_IN_CHANNELS, pretrained=False)
            input_feats_dim, fc1_dim = 2048, 1024
        else:
            raise ValueError('NUM_RESNET_LAYERS must be 18 or 50')
        self.input_feats_dim = input_feats_dim
        self.activation = nn.ELU()
        self.fc1 = nn.Linear(input_feats_dim, fc1_dim)
        self.fc_shape = nn.Linear(fc1_dim, self.num_shape_params * 2)
        self.fc_glob = nn.Linear(fc1_dim, self.num_glob_params)
        self.fc_cam = nn.Linear(fc1_dim, self.num_cam_params)
        nf = model_cfg.NORM_FLOW
        if nf.TRANSFORM_TYPE != 'spline_coupling' or nf.PERMUTE_TYPE != 'permute':
            raise NotImplementedError('only the default spline_coupling + permute flow is implemented '
                                      '(configs/humaniflow_config.py:17,20; other variants are out of scope, SURVEY.md 2)')
        self.fc_input_shape_glob_cam_feats = nn.Linear(input_feats_dim + self.num_shape_params + 9 + self.num_cam_params,
                                                       model_cfg.INPUT_SHAPE_GLOB_CAM_FEATS_DIM)
        self.fc_flow_context = nn.ModuleList()
        self.pose_so3flow_transform_modules = nn.ModuleList()
        self.pose_SO3flow_dists = []
        for bodypart in range(self.num_bodyparts):
            na = len(self.ancestors_dict[bodypart])
            self.fc_flow_context.append(nn.Linear(model_cfg.INPUT_SHAPE_GLOB_CAM_FEATS_DIM + na * 9, nf.CONTEXT_DIM))
            for _ in range(nf.NUM_TRANSFORMS):
                self.pose_so3flow_transform_modules.append(ConditionalSplineCoupling(
                    3, nf.CONTEXT_DIM, nf.TRANSFORM_NN_HIDDEN_DIMS, nf.NUM_SPLINE_SEGMENTS, nf.COMPACT_SUPPORT_RADIUS))
            self.pose_SO3flow_dists.append(_FlowDistHandle())
        self._flow = None
        self._packed = None
        self._packed_version = None
        self._index_cache = {}
        self._flow_ws = None

    # ------------------------------------------------------------------ packing
    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._drop()
        return out

    def _drop(self):
        if getattr(self, '_flow', None) is not None:
            try:
                _lib.load().hf_flow_destroy(self._flow)
            except Exception:
                pass
        self._flow = None
        self._packed = None
        self._packed_version = None

    def __del__(self):
        try:
            self._drop()
        except Exception:
            pass

    def _head_params(self):
        mods = [self.fc1, self.fc_shape, self.fc_glob, self.fc_cam, self.fc_input_shape_glob_cam_feats,
                self.fc_flow_context, self.pose_so3flow_transform_modules]
        return [p for m in mods for p in m.parameters()]

    def _ensure_packed(self, device):
        ver = tuple((p.data_ptr(), p._version) for p in self._head_params() + [self.init_glob, self.init_cam])
        if self._flow is not None and ver == self._packed_version:
            return
        self._drop()
        lib = _lib.load()
        nf = self.cfg.NORM_FLOW
        cpu = lambda t: t.detach().to('cpu', torch.float32).contiguous()
        F, nb = self.input_feats_dim, self.num_shape_params
        Wimg = cpu(self.fc_input_shape_glob_cam_feats.weight)
        keep = []
        anc, offs = [], [0]
        for j in range(self.num_bodyparts):
            anc += list(self.ancestors_dict[j])
            offs.append(len(anc))
        anc_t = torch.tensor(anc if anc else [0], dtype=torch.int32)
        offs_t = torch.tensor(offs, dtype=torch.int32)
        beta_w = Wimg[:, F:F + nb].contiguous()
        ctx_w = [cpu(m.weight) for m in self.fc_flow_context]
        ctx_b = [cpu(m.bias) for m in self.fc_flow_context]
        nn_w = [cpu(l.weight) for c in self.pose_so3flow_transform_modules for l in c.nn.layers]
        nn_b = [cpu(l.bias) for c in self.pose_so3flow_transform_modules for l in c.nn.layers]
        keep += [anc_t, offs_t, beta_w] + ctx_w + ctx_b + nn_w + nn_b
        arr = lambda ts: (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
        cfg = _lib.FlowConfig(self.num_bodyparts, self.cfg.INPUT_SHAPE_GLOB_CAM_FEATS_DIM, nf.CONTEXT_DIM, nf.NUM_TRANSFORMS,
                              (ctypes.c_int * 3)(*nf.TRANSFORM_NN_HIDDEN_DIMS), nf.NUM_SPLINE_SEGMENTS, nb,
                              float(nf.COMPACT_SUPPORT_RADIUS), float(nf.BASE_DIST_STD))
        h = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.hf_flow_create(ctypes.byref(h), ctypes.byref(cfg), _lib.ptr(anc_t), _lib.ptr(offs_t), _lib.ptr(beta_w),
                                          arr(ctx_w), arr(ctx_b), arr(nn_w), arr(nn_b)))
        self._flow = h
        dev = lambda t: t.to(device).contiguous()
        self._packed = {
            # heads as one (2nb+9, fc1) matrix: [shape | glob | cam]
            'heads_w': dev(torch.cat([cpu(self.fc_shape.weight), cpu(self.fc_glob.weight), cpu(self.fc_cam.weight)], 0)),
            'heads_b': dev(torch.cat([cpu(self.fc_shape.bias), cpu(self.fc_glob.bias), cpu(self.fc_cam.bias)], 0)),
            # image-level Linear split by input block (feats | beta | glob | cam) so every matrix is 16-byte aligned
            'img_wf': dev(Wimg[:, :F]), 'img_wg': dev(Wimg[:, F + nb:F + nb + 9]), 'img_wc': dev(Wimg[:, F + nb + 9:]),
            'img_b': dev(cpu(self.fc_input_shape_glob_cam_feats.bias)),
            'fc1_w': dev(cpu(self.fc1.weight)), 'fc1_b': dev(cpu(self.fc1.bias)),
            'init_glob': dev(cpu(self.init_glob).view(-1)), 'init_cam': dev(cpu(self.init_cam).view(-1)),
        }
        self._packed_version = ver

    def _linear(self, x, W, b, out, K, O, act=0, accumulate=0, w_offset=0):
        lib = _lib.load()
        wp = ctypes.c_void_p(W.data_ptr() + 4 * w_offset)
        M = x.shape[0]
        nbytes = lib.hf_linear_workspace_bytes(M, K, O)         # partial sums of the K-sliced layers (one buffer per model: single stream)
        ws = getattr(self, '_lin_ws', None)
        if nbytes and (ws is None or ws.numel() < nbytes or ws.device != x.device):
            ws = torch.empty(nbytes, device=x.device, dtype=torch.uint8)
            self._lin_ws = ws
        _lib.check(lib.hf_linear_ws(_lib.ptr(x), x.stride(0), wp, W.stride(0), _lib.ptr(b), _lib.ptr(out), out.stride(0),
                                    M, K, O, act, accumulate, _lib.ptr(ws) if nbytes else None, ws.numel() if nbytes else 0, _lib.stream()))

    def _img_base(self, input_feats, glob_R, cam):
        """W[:, feats|glob|cam] . [feats, vec(glob_R), cam] + b : the beta-independent part of
        models/humaniflow_model.py:133-148 (the beta term and the ELU are applied per row in the flow kernel)."""
        P = self._packed
        B, F, nb = input_feats.shape[0], self.input_feats_dim, self.num_shape_params
        D = self.cfg.INPUT_SHAPE_GLOB_CAM_FEATS_DIM
        base = torch.empty(B, D, device=input_feats.device, dtype=torch.float32)
        self._linear(input_feats, P['img_wf'], P['img_b'], base, F, D)
        g = _lib.f32c(glob_R).reshape(B, 9)
        self._linear(g, P['img_wg'], None, base, 9, D, accumulate=1)
        self._linear(cam, P['img_wc'], None, base, 3, D, accumulate=1)
        return base

    def _img_index(self, B, N, with_pe, device):
        key = (B, N, with_pe, str(device))
        if key not in self._index_cache:
            idx = torch.arange(B, device=device, dtype=torch.int32).repeat_interleave(N)
            if with_pe:
                idx = torch.cat([idx, torch.arange(B, device=device, dtype=torch.int32)])
            self._index_cache[key] = idx.contiguous()
        return self._index_cache[key]

    # ------------------------------------------------------------------ forward
    def forward(self, *args, **kwargs):
        """See models/humaniflow_model.py:188-340 for the argument meaning.  The network itself runs without an autograd graph
        (inference kernels); when autograd is recording, the teacher-forced contexts of ``compute_for_loglik=True`` stay
        differentiable w.r.t. shape_for_loglik / pose_R_for_loglik / glob_R_for_loglik (and input_feats), and the returned
        conditioned distributions' ``log_prob`` w.r.t. their contexts and targets: the pose-prior term of a fitting loop
        (optimise/optimise_humaniflow.py:96-114) can be back-propagated to the pose, shape and global rotation."""
        return self._forward(torch.is_grad_enabled(), *args, **kwargs)

    @torch.no_grad()
    def _forward(self, grad_on, input, compute_point_est=True, num_samples=0, use_shape_mode_for_samples=False,
                 compute_for_loglik=False, shape_for_loglik=None, pose_R_for_loglik=None, glob_R_for_loglik=None,
                 input_feats=None, grad_for_pose_point_est=False, return_input_feats=False,
                 return_input_feats_only=False, *, base_noise=None, shape_eps=None):
        _lib.require_cuda('HumaniflowModel.forward')
        lib = _lib.load()
        if input_feats is None:
            input_feats = self.image_encoder(input)                                   # :215-216
        if return_input_feats_only:
            return {'input_feats': input_feats}
        if not input_feats.is_cuda:
            raise RuntimeError('humaniflow_b200: input_feats must be a CUDA tensor (no CPU fallback)')
        dev = input_feats.device
        input_feats = _lib.f32c(input_feats)
        B = input_feats.shape[0]
        if compute_for_loglik:                                                        # :224-230
            assert pose_R_for_loglik is not None and pose_R_for_loglik.shape[0] == B
            assert shape_for_loglik is not None and shape_for_loglik.shape[0] == B
            assert glob_R_for_loglik is not None and glob_R_for_loglik.shape[0] == B
        N, nb, J = int(num_samples), self.num_shape_params, self.num_bodyparts
        nf = self.cfg.NORM_FLOW
        with torch.cuda.device(dev):
            self._ensure_packed(dev)
            P = self._packed
            st = _lib.stream()
            # heads (:232-258)
            x = torch.empty(B, self.fc1.out_features, device=dev, dtype=torch.float32)
            self._linear(input_feats, P['fc1_w'], P['fc1_b'], x, self.input_feats_dim, x.shape[1], act=1)
            heads = torch.empty(B, 2 * nb + 9, device=dev, dtype=torch.float32)
            self._linear(x, P['heads_w'], P['heads_b'], heads, x.shape[1], heads.shape[1])
            cam = torch.empty(B, 3, device=dev, dtype=torch.float32)
            glob6 = torch.empty(B, 6, device=dev, dtype=torch.float32)
            shape_rows = torch.empty(B * N + B, nb, device=dev, dtype=torch.float32)
            eps = None
            if N > 0 and not use_shape_mode_for_samples:
                eps = torch.randn(B, N, nb, device=dev) if shape_eps is None else _lib.f32c(shape_eps, dev)
                assert eps.shape == (B, N, nb)
            glob_R = torch.empty(B, 3, 3, device=dev, dtype=torch.float32)
            shape_std = torch.empty(B, nb, device=dev, dtype=torch.float32)
            _lib.check(lib.hf_heads_finish(_lib.ptr(heads), _lib.ptr(P['init_glob']), _lib.ptr(P['init_cam']), _lib.ptr(eps),
                                           B, N, nb, _lib.ptr(cam), _lib.ptr(glob6), _lib.ptr(shape_rows), _lib.ptr(glob_R),
                                           _lib.ptr(shape_std), st))
            shape_mode, shape_log_std = heads[:, :nb], heads[:, nb:2 * nb]
            shape_dist = Normal(loc=shape_mode, scale=shape_std, validate_args=False)
            out = {'cam_wp': cam, 'glob_rotmat': glob_R, 'shape_mode': shape_mode, 'shape_log_std': shape_log_std,
                   'shape_dist_for_loglik': shape_dist}
            # pose: one launch walks all 23 joints for the N samples and the point estimate (:263-311)
            if compute_point_est or N > 0:
                base = self._img_base(input_feats, glob_R, cam)
                Rn = B * N
                R = Rn + (B if compute_point_est else 0)
                rows = shape_rows if compute_point_est else shape_rows[:Rn]
                idx = self._img_index(B, N, compute_point_est, dev)
                noise = None
                if N > 0:
                    noise = (torch.randn(B, N, J, 3, device=dev) * float(nf.BASE_DIST_STD)) if base_noise is None \
                        else _lib.f32c(base_noise, dev)
                    assert noise.shape == (B, N, J, 3)
                rot = torch.empty(R, J, 3, 3, device=dev, dtype=torch.float32)
                aa = torch.empty(B, J, 3, device=dev, dtype=torch.float32) if compute_point_est else None
                nbytes = lib.hf_flow_workspace_bytes(self._flow, R)
                if self._flow_ws is None or self._flow_ws.numel() < nbytes or self._flow_ws.device != dev:
                    self._flow_ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
                _lib.check(lib.hf_flow_sample(self._flow, _lib.ptr(base), _lib.ptr(rows), _lib.ptr(idx), _lib.ptr(noise), R, Rn,
                                              _lib.ptr(rot), _lib.ptr(aa), _lib.ptr(self._flow_ws), self._flow_ws.numel(), st))
                if compute_point_est:
                    out['pose_axisangle_point_est'] = aa
                    out['pose_rotmats_point_est'] = rot[Rn:]
                if N > 0:
                    out['pose_rotmats_samples'] = rot[:Rn].view(B, N, J, 3, 3)
                    out['shape_samples'] = shape_rows[:Rn].view(B, N, nb)
            # teacher-forced contexts for the log-likelihood (:277-283, :314-320)
            if compute_for_loglik:
                wants_grad = grad_on and any(torch.is_tensor(t) and t.requires_grad
                                             for t in (shape_for_loglik, pose_R_for_loglik, glob_R_for_loglik, input_feats))
                to_dev = lambda t: t.to(dev)
                with torch.set_grad_enabled(wants_grad):
                    ctx = _FlowContextFunction.apply(self, input_feats, cam, to_dev(shape_for_loglik), to_dev(pose_R_for_loglik),
                                                     to_dev(glob_R_for_loglik))
                out['conditioned_pose_so3flow_dists_for_loglik'] = [ConditionedSO3FlowDist(self, j, ctx, False) for j in range(J)]
                out['conditioned_pose_SO3flow_dists_for_loglik'] = [ConditionedSO3FlowDist(self, j, ctx, True) for j in range(J)]
                out['flow_contexts_for_loglik'] = ctx
        if return_input_feats:
            out['input_feats'] = input_feats
        return out

    def pose_log_prob(self, ctx, pose_R):
        """All joints at once: ctx (B,23,64) from ``flow_contexts_for_loglik``, pose_R (B,23,3,3) -> (B,23) fp32
        (= stacking ``dist_j.log_prob(pose_R[:, j].double())`` as losses/humaniflow_loss.py:25-35 does)."""
        lib = _lib.load()
        B, J = ctx.shape[0], self.num_bodyparts
        if torch.is_grad_enabled() and (ctx.requires_grad or pose_R.requires_grad):
            return _FlowLogProbFunction.apply(self, 0, J, True, ctx, pose_R.to(ctx.device))
        ctx = _lib.f32c(ctx)
        v = pose_R.detach().to(ctx.device, torch.float64).contiguous()
        out = torch.empty(B, J, device=ctx.device, dtype=torch.float32)
        with torch.cuda.device(ctx.device):
            _lib.check(lib.hf_flow_log_prob(self._flow, _lib.ptr(ctx), ctx.shape[1] * ctx.shape[2], 0, J, _lib.ptr(v), B,
                                            _lib.ptr(out), _lib.stream()))
        return out
