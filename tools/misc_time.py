"""CUDA-event timings of the kernels around the hot path at the bench sizes (B=32, N=100): proxy representation, per-sample error
metrics, sample statistics, heads."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import humaniflow_b200 as hb  # noqa: E402
from humaniflow_b200 import _lib  # noqa: E402
from humaniflow_b200.metrics import pointset_errors, sample_stats  # noqa: E402
from humaniflow_b200.proxy_rep import build_proxy_representation  # noqa: E402
from humaniflow_b200.synthetic import SMPL_PARENTS  # noqa: E402


def timeit(name, fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    print('%-40s %8.1f us' % (name, a.elapsed_time(b) / iters * 1e3))


B, N, V = 32, 100, 6890
torch.manual_seed(0)
cfg = hb.get_model_cfg_defaults()
cfg.NUM_RESNET_LAYERS = 50
m = hb.HumaniflowModel('cuda', cfg, SMPL_PARENTS).eval().cuda()
rgb = torch.rand(B, 3, 256, 256, device='cuda')
j2d = torch.rand(B, 17, 2, device='cuda') * 256
verts = torch.randn(B, N, V, 3, device='cuda')
tgt = torch.randn(B, V, 3, device='cuda')
joints = torch.randn(B, N, 90, 3, device='cuda')
feats = torch.randn(B, 2048, device='cuda').abs()
x = torch.rand(B, 18, 256, 256, device='cuda')
m.image_encoder(x)
timeit('proxy_rep (fp32 NCHW out)', lambda: build_proxy_representation(rgb, j2d))
timeit('proxy_rep staged (bf16 NHWC into the encoder)', lambda: build_proxy_representation(rgb, j2d, encoder=m.image_encoder))
timeit('pointset_errors (32,100,6890)', lambda: pointset_errors(verts, tgt))
timeit('sample_stats joints (32,100,90)', lambda: sample_stats(joints))
timeit('sample_stats verts (32,100,6890)', lambda: sample_stats(verts))
timeit('model heads + point estimate only (N=0)', lambda: m(None, input_feats=feats))
lib = _lib.load()
y = torch.empty(B, 1024, device='cuda')
W = torch.randn(1024, 2048, device='cuda')
b = torch.randn(1024, device='cuda')
timeit('hf_linear 32 x 2048 -> 1024', lambda: _lib.check(lib.hf_linear(_lib.ptr(feats), 2048, _lib.ptr(W), 2048, _lib.ptr(b), _lib.ptr(y), 1024, B, 2048, 1024, 1, 0, _lib.stream())))
W2 = torch.randn(256, 2048, device='cuda')
y2 = torch.empty(B, 256, device='cuda')
timeit('hf_linear 32 x 2048 -> 256', lambda: _lib.check(lib.hf_linear(_lib.ptr(feats), 2048, _lib.ptr(W2), 2048, None, _lib.ptr(y2), 256, B, 2048, 256, 0, 0, _lib.stream())))
x3 = torch.randn(B, 1024, device='cuda')
W3 = torch.randn(29, 1024, device='cuda')
y3 = torch.empty(B, 29, device='cuda')
timeit('hf_linear 32 x 1024 -> 29', lambda: _lib.check(lib.hf_linear(_lib.ptr(x3), 1024, _lib.ptr(W3), 1024, None, _lib.ptr(y3), 29, B, 1024, 29, 0, 0, _lib.stream())))
