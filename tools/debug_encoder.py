"""Layer-by-layer bring-up of the encoder on the GPU: after every op compare the activation of the tcgen05 path
and of the SIMT path with the CPU restatement of the bf16 contract (tests/util.py)."""
import ctypes
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
from humaniflow_b200 import _lib  # noqa: E402
from util import bf16r, conv_bf16_ref, make_model  # noqa: E402


def cpu_activations(sd, x, layers):
    """list of NHWC activations after every op of the program (stem, maxpool, then each conv in program order)."""
    s = {k[len('image_encoder.'):]: v for k, v in sd.items() if k.startswith('image_encoder.')}
    acts = []

    def fold(conv, bn):
        w = s[conv + '.weight'].double()
        scale = s[bn + '.weight'].double() / torch.sqrt(s[bn + '.running_var'].double() + 1e-5)
        b = s[bn + '.bias'].double() - s[bn + '.running_mean'].double() * scale
        return bf16r((w * scale[:, None, None, None]).float()).permute(0, 2, 3, 1), b.float()

    def conv(x, c, b, stride, pad, relu, res=None):
        w, bb = fold(c, b)
        y = conv_bf16_ref(x, w, bb, res, stride, pad, relu)
        acts.append(y)
        return y
    kind, counts = ('basic', [2, 2, 2, 2]) if layers == 18 else ('bottleneck', [3, 4, 6, 3])
    x = bf16r(x).permute(0, 2, 3, 1)
    x = conv(x, 'conv1', 'bn1', 2, 3, True)
    x = F.max_pool2d(x.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    acts.append(x)
    for li, n in enumerate(counts, start=1):
        for bi in range(n):
            p = 'layer%d.%d.' % (li, bi)
            stride = 2 if (li > 1 and bi == 0) else 1
            has_ds = p + 'downsample.0.weight' in s
            if kind == 'basic':
                o = conv(x, p + 'conv1', p + 'bn1', stride, 1, True)
                idn = conv(x, p + 'downsample.0', p + 'downsample.1', stride, 0, False) if has_ds else x
                x = conv(o, p + 'conv2', p + 'bn2', 1, 1, True, res=idn)
            else:
                o = conv(x, p + 'conv1', p + 'bn1', 1, 0, True)
                o = conv(o, p + 'conv2', p + 'bn2', stride, 1, True)
                idn = conv(x, p + 'downsample.0', p + 'downsample.1', stride, 0, False) if has_ds else x
                x = conv(o, p + 'conv3', p + 'bn3', 1, 0, True, res=idn)
    return acts


def main(layers=18, size=64, B=2):
    lib = _lib.load()
    m, sd, cfg = make_model(layers, seed=20 + layers)
    enc = m.image_encoder.cuda()
    x = torch.rand(B, 18, size, size, generator=torch.Generator().manual_seed(5))
    acts = cpu_activations(sd, x, layers)
    xd = x.cuda()
    enc(xd)   # builds the handle
    nops = len(acts)
    for impl in (0, 1):
        enc.set_impl(impl)
        print('impl', impl)
        for i in range(nops):
            ref = acts[i]
            out = torch.empty(ref.shape, dtype=torch.int16, device='cuda')
            dims = (ctypes.c_int * 3)()
            feats = torch.empty(B, enc.feat_dim, device='cuda')
            _lib.check(lib.hf_encoder_debug_op_output(enc._enc, i, _lib.ptr(xd), B, size, size, _lib.ptr(enc._ws), enc._ws.numel(),
                                                      _lib.ptr(out), out.numel() * 2, dims, _lib.ptr(feats), _lib.stream()))
            torch.cuda.synchronize()
            got = out.view(torch.bfloat16).float().cpu()
            rel = ((got - ref).norm() / ref.norm().clamp_min(1e-9)).item()
            bad = (~torch.isfinite(got)).sum().item()
            print('  op %2d dims %s ref %s rel %.3e nonfinite %d max|got| %.3e' % (i, list(dims), list(ref.shape[1:]), rel, bad, got.abs().max().item()))
            if rel > 0.05 and i < 3:
                d = (got - ref).abs()
                idx = torch.nonzero(d > 0.05 * ref.abs().max())[:8]
                print('     first bad idx (b,h,w,c):', idx.tolist())
                print('     got', got[0, :2, :3, :4].tolist())
                print('     ref', ref[0, :2, :3, :4].tolist())
    enc.set_impl(0)


if __name__ == '__main__':
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 18, int(sys.argv[2]) if len(sys.argv) > 2 else 64)
