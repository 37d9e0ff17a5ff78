"""Shared helpers for the parity tests (test-side only)."""
import math
import os

import numpy as np
import torch
import torch.nn.functional as F

import humaniflow_b200 as hb
from humaniflow_b200.synthetic import SMPL_PARENTS, synthetic_smpl_data

RADIUS = 1.5 * math.pi
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def real_regressors():
    """The reference's three shipped regressors (model_files/*.npy), from the committed sparse fixture."""
    g = np.load(os.path.join(GOLDEN, 'regressors_sparse.npz'))
    out = {}
    for key, name in (('extra', 'J_regressor_extra'), ('cocoplus', 'J_regressor_cocoplus'), ('h36m', 'J_regressor_h36m')):
        a = np.zeros(tuple(g[key + '_shape']), dtype=np.float64)
        a[g[key + '_row'], g[key + '_col']] = g[key + '_val']
        out[name] = a
    return out


_SMPL_CACHE = {}


def smpl_data(real_regs=True):
    if real_regs not in _SMPL_CACHE:
        _SMPL_CACHE[real_regs] = synthetic_smpl_data(seed=0, regressors=real_regressors() if real_regs else None)
    return _SMPL_CACHE[real_regs]


def make_model(layers=18, seed=0, flow_scale=1.0, bn_stats=True, in_channels=18, res_gain=None):
    """HumaniflowModel with torch-default init (seeded); optional scaling of the flow weights to make the
    splines less trivial, and non-trivial BatchNorm statistics so eval-mode BN is not a no-op.  ``res_gain`` scales the last
    BatchNorm of every residual block (the usual small-gamma initialisation of trained ResNets): without it the random-init
    ResNet-50 features grow to O(100) through 16 un-normalised residual additions and the heads saturate."""
    torch.manual_seed(seed)
    cfg = hb.get_model_cfg_defaults()
    cfg.NUM_RESNET_LAYERS = layers
    cfg.NUM_IN_CHANNELS = in_channels
    m = hb.HumaniflowModel('cpu', cfg, SMPL_PARENTS)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for k, v in m.state_dict().items():
            if k.startswith('pose_so3flow') or k.startswith('fc_flow_context'):
                v.mul_(flow_scale)
            if bn_stats and k.endswith('running_mean'):
                v.copy_(torch.randn(v.shape, generator=g) * 0.1)
            if bn_stats and k.endswith('running_var'):
                v.copy_(torch.rand(v.shape, generator=g) + 0.5)
            if bn_stats and ('.bn' in k or 'downsample.1' in k or k.startswith('image_encoder.bn1')) and k.endswith('.weight'):
                v.copy_(torch.rand(v.shape, generator=g) * 0.5 + 0.75)
            if bn_stats and ('.bn' in k or 'downsample.1' in k or k.startswith('image_encoder.bn1')) and k.endswith('.bias'):
                v.copy_(torch.randn(v.shape, generator=g) * 0.1)
            if res_gain is not None and (('.bn3.' in k) or (layers == 18 and '.bn2.' in k)) and (k.endswith('.weight') or k.endswith('.bias')):
                v.mul_(res_gain)
    m.eval()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    return m, sd, cfg


def bf16r(t):
    return t.to(torch.bfloat16).to(torch.float32)


def conv_bf16_ref(x_nhwc, w_ohwi, bias, res, stride, pad, relu):
    """Reference for one fused conv on bf16-rounded operands: fp32 conv, +bias, +res, relu, round to bf16.
    x (B,H,W,Ci), w (Co,k,k,Ci), res (B,Ho,Wo,Co) or None; all fp32 tensors holding bf16-representable values."""
    y = F.conv2d(x_nhwc.permute(0, 3, 1, 2), w_ohwi.permute(0, 3, 1, 2), stride=stride, padding=pad)
    y = y.permute(0, 2, 3, 1) + bias
    if res is not None:
        y = y + res
    if relu:
        y = F.relu(y)
    return bf16r(y)


def resnet_bf16emu(sd, x, layers, prefix='image_encoder.'):
    """The encoder's numerics contract restated on the CPU: BatchNorm folded (fp64) into bf16 weights + fp32 bias,
    bf16 activations between layers, fp32 accumulation.  Used to check the CUDA encoder tightly; the plain fp32
    oracle (oracle.resnet) bounds the bf16 contract itself."""
    s = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}

    def fold(conv, bn):
        w = s[conv + '.weight'].double()
        scale = s[bn + '.weight'].double() / torch.sqrt(s[bn + '.running_var'].double() + 1e-5)
        b = s[bn + '.bias'].double() - s[bn + '.running_mean'].double() * scale
        return bf16r((w * scale[:, None, None, None]).float()).permute(0, 2, 3, 1), b.float()

    def conv(x, cname, bname, stride, pad, relu, res=None):
        w, b = fold(cname, bname)
        return conv_bf16_ref(x, w, b, res, stride, pad, relu)

    kind, counts = ('basic', [2, 2, 2, 2]) if layers == 18 else ('bottleneck', [3, 4, 6, 3])
    x = bf16r(x).permute(0, 2, 3, 1)
    x = conv(x, 'conv1', 'bn1', 2, 3, True)
    x = F.max_pool2d(x.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    for li, n in enumerate(counts, start=1):
        for bi in range(n):
            p = 'layer%d.%d.' % (li, bi)
            stride = 2 if (li > 1 and bi == 0) else 1
            idn = x
            if p + 'downsample.0.weight' in s:
                idn = conv(x, p + 'downsample.0', p + 'downsample.1', stride, 0, False)
            if kind == 'basic':
                o = conv(x, p + 'conv1', p + 'bn1', stride, 1, True)
                x = conv(o, p + 'conv2', p + 'bn2', 1, 1, True, res=idn)
            else:
                o = conv(x, p + 'conv1', p + 'bn1', 1, 0, True)
                o = conv(o, p + 'conv2', p + 'bn2', stride, 1, True)
                x = conv(o, p + 'conv3', p + 'bn3', 1, 0, True, res=idn)
    return x.mean(dim=(1, 2))


def special_rotations():
    """Target rotations for log_prob incl. the reference's delicate cases:
    theta < 1e-6, theta ~ pi/2 (second pre-image enters the support), |pi - theta| < 1e-2."""
    from oracle import so3
    rs = np.random.RandomState(5)
    axes = rs.standard_normal((40, 3))
    axes /= np.linalg.norm(axes, axis=1, keepdims=True)
    ang = np.concatenate([rs.uniform(0.05, 3.0, 24), [0.0, 1e-9, 1e-7, 5e-7, math.pi / 2 - 1e-3, math.pi / 2 + 1e-3,
                          math.pi / 2, 2.0, 2.5, 3.0, math.pi - 5e-3, math.pi - 1e-3, math.pi - 1e-4, math.pi - 1e-6,
                          math.pi - 9e-3, 3.1]])
    return so3.so3_exp(torch.tensor(axes * ang[:, None], dtype=torch.float64))
