"""ctypes binding of include/humaniflow_b200.h (the C-ABI drop-in boundary).

There is NO CPU fallback: if the CUDA library has not been built, or no CUDA device is present when a
compute entry point is called, this raises.  Build with ``python -m humaniflow_b200.build``.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('HF_LIB_PATH') or os.path.join(_HERE, 'lib', 'libhumaniflow_b200.so')   # HF_LIB_PATH: A/B builds of the same ABI (tools/)

c_void_p, c_int, c_size_t, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_float


class FlowConfig(ctypes.Structure):
    _fields_ = [('num_joints', c_int), ('feats_dim', c_int), ('context_dim', c_int), ('num_transforms', c_int),
                ('hidden', c_int * 3), ('num_bins', c_int), ('num_betas', c_int), ('radius', c_float),
                ('base_std', c_float)]


class EncOp(ctypes.Structure):
    _fields_ = [('kind', c_int), ('src', c_int), ('dst', c_int), ('res', c_int), ('cin', c_int), ('cout', c_int),
                ('ksize', c_int), ('stride', c_int), ('pad', c_int), ('relu', c_int), ('weight_index', c_int),
                ('src2', c_int), ('cin2', c_int), ('stride2', c_int)]

    def __init__(self, kind, src, dst, res, cin, cout, ksize, stride, pad, relu, weight_index, src2=-1, cin2=0, stride2=1):
        super().__init__(kind, src, dst, res, cin, cout, ksize, stride, pad, relu, weight_index, src2, cin2, stride2)


OP_CONV, OP_MAXPOOL, OP_AVGPOOL = 0, 1, 2

# name -> (restype, argtypes); every symbol include/humaniflow_b200.h declares
SIGNATURES = {
    'hf_version': (c_int, []),
    'hf_last_error': (ctypes.c_char_p, []),
    'hf_launch_count': (ctypes.c_longlong, []),
    'hf_smpl_create': (c_int, [ctypes.POINTER(c_void_p), c_int, c_int, c_int] + [c_void_p] * 7 + [c_int, c_void_p, c_int]),
    'hf_smpl_destroy': (None, [c_void_p]),
    'hf_smpl_num_joints_out': (c_int, [c_void_p]),
    'hf_lbs_workspace_bytes': (c_size_t, [c_void_p, c_int]),
    'hf_lbs_set_impl': (c_int, [c_void_p, c_int]),
    'hf_lbs_forward': (c_int, [c_void_p] * 7 + [c_size_t, c_int, c_void_p]),
    'hf_lbs_backward_workspace_bytes': (c_size_t, [c_void_p, c_int]),
    'hf_lbs_backward': (c_int, [c_void_p] * 7 + [c_void_p, c_size_t, c_int, c_void_p]),
    'hf_lbs_forward_split': (c_int, [c_void_p] * 4 + [c_int] + [c_void_p] * 4 + [c_size_t, c_int, c_void_p]),
    'hf_rodrigues': (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    'hf_lbs_tpose': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    'hf_smpl_dims': (c_int, [c_void_p] + [ctypes.POINTER(c_int)] * 5),
    'hf_vertex_variance': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'hf_pointset_errors': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    'hf_pointset_errors_workspace_bytes': (c_size_t, [c_int, c_int]),
    'hf_pointset_errors_ws': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    'hf_sample_stats': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'hf_samples_reduce': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    'hf_proxy_rep': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, ctypes.c_float, c_int, ctypes.c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    'hf_proxy_rep_staged': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, ctypes.c_float, c_int, ctypes.c_float, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'hf_project_joints2d': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, ctypes.c_float, c_void_p, c_void_p]),
    'hf_flow_create': (c_int, [ctypes.POINTER(c_void_p), ctypes.POINTER(FlowConfig)] + [c_void_p] * 7),
    'hf_flow_destroy': (None, [c_void_p]),
    'hf_flow_workspace_bytes': (c_size_t, [c_void_p, c_int]),
    'hf_flow_sample': (c_int, [c_void_p] * 5 + [c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'hf_flow_context': (c_int, [c_void_p] * 5 + [c_int, c_void_p, c_void_p]),
    'hf_flow_log_prob': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    'hf_flow_algebra_log_prob': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    'hf_flow_log_prob_backward': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    'hf_flow_algebra_log_prob_backward': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    'hf_linear': (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'hf_linear_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'hf_linear_ws': (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    'hf_heads_finish': (c_int, [c_void_p] * 4 + [c_int, c_int, c_int] + [c_void_p] * 6),
    'hf_rot6d_to_rotmat': (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    'hf_encoder_create': (c_int, [ctypes.POINTER(c_void_p), ctypes.POINTER(EncOp), c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int]),
    'hf_encoder_destroy': (None, [c_void_p]),
    'hf_encoder_workspace_bytes': (c_size_t, [c_void_p, c_int, c_int, c_int]),
    'hf_encoder_forward': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    'hf_encoder_forward_bf16': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    'hf_encoder_forward_staged': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    'hf_encoder_stem_input': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p, ctypes.POINTER(c_void_p), ctypes.POINTER(c_int)]),
    'hf_encoder_invalidate': (c_int, [c_void_p]),
    'hf_encoder_stem_channels': (c_int, [c_void_p]),
    'hf_encoder_set_impl': (c_int, [c_void_p, c_int]),
    'hf_encoder_debug_op_output': (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p]),
    'hf_debug_conv_stamps': (c_int, [c_void_p]),
    'hf_conv2d_nhwc': (c_int, [c_void_p] * 5 + [c_int] * 10 + [c_void_p]),
    'hf_conv2d_nhwc_branch': (c_int, [c_void_p] * 5 + [c_int] * 9 + [c_void_p] + [c_int] * 5 + [c_void_p]),
}

_lib = None


def load():
    """Load the shared library (no GPU needed for loading / symbol lookup)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError('humaniflow_b200: CUDA library %s is missing; run `python -m humaniflow_b200.build` '
                               '(there is no CPU fallback)' % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def require_cuda(what):
    if not torch.cuda.is_available():
        raise RuntimeError('humaniflow_b200.%s needs a CUDA device (sm_100a); there is no CPU fallback' % what)


def check(rc):
    if rc != 0:
        raise RuntimeError('humaniflow_b200 error %d: %s' % (rc, load().hf_last_error().decode()))


def ptr(t):
    return None if t is None else c_void_p(t.data_ptr())


def stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def f32c(t, device=None):
    """contiguous fp32 view/copy of a tensor (optionally on `device`)."""
    t = t.detach()
    if device is not None:
        t = t.to(device)
    return t.to(torch.float32).contiguous()


def launch_count():
    return int(load().hf_launch_count())
