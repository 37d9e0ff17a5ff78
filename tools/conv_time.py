"""Time single convolutions (ResNet-50 layer shapes at B=32) through hf_conv2d_nhwc with CUDA events; HF_CONV_DBG /
HF_CONV_PAIR select the timing experiment (see encoder.cu)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from humaniflow_b200 import _lib  # noqa: E402

lib = _lib.load()
SHAPES = {   # name: B, H, W, Cin, Cout, k, stride, pad, res
    'l1c2': (32, 64, 64, 64, 64, 3, 1, 1, 0),
    'l1c3': (32, 64, 64, 64, 256, 1, 1, 0, 1),
    'l2c2': (32, 32, 32, 128, 128, 3, 1, 1, 0),
    'l3c1': (32, 16, 16, 1024, 256, 1, 1, 0, 0),
    'l3c2': (32, 16, 16, 256, 256, 3, 1, 1, 0),
    'l3c3': (32, 16, 16, 256, 1024, 1, 1, 0, 1),
    'l4c2': (32, 8, 8, 512, 512, 3, 1, 1, 0),
    'e64': (32, 64, 64, 128, 64, 3, 1, 1, 0),         # BN=64, even k-block count, many tiles per CTA
    'so64': (2, 32, 32, 64, 64, 3, 1, 1, 0),          # BN=64, odd k-block count, one tile per CTA
    'se64': (2, 32, 32, 128, 64, 3, 1, 1, 0),
    'o128': (32, 32, 32, 64, 128, 3, 1, 1, 0),        # BN=128, odd k-block count (9), two tiles per CTA
    'l1c1': (32, 64, 64, 256, 64, 1, 1, 0, 0),
}
names = sys.argv[1:] or list(SHAPES)
for name in names:
    B, H, W, Ci, Co, k, s, p, res = SHAPES[name]
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    x = torch.randn(B, H, W, Ci, device='cuda').to(torch.bfloat16)
    w = (torch.randn(Co, k, k, Ci, device='cuda') * 0.05).to(torch.bfloat16)
    b = torch.zeros(Co, device='cuda')
    r = torch.randn(B, Ho, Wo, Co, device='cuda').to(torch.bfloat16) if res else None
    y = torch.empty(B, Ho, Wo, Co, device='cuda', dtype=torch.bfloat16)
    # scratch kernel between timed launches so that the conv never overlaps itself through PDL
    def run():
        _lib.check(lib.hf_conv2d_nhwc(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(r), _lib.ptr(y), B, H, W, Ci, Co, k, s, p, 1, 0, _lib.stream()))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); e.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(e) * 1e3)
    ts.sort()
    nkb = k * k * Ci // 64
    tiles = (B * Ho * Wo // 128) * (Co // (128 if Co % 128 == 0 else 64))
    print('%s dbg=%s pair=%s: median %.1f us  (tiles %d, k-blocks/tile %d)' % (name, os.environ.get('HF_CONV_DBG', '0'), os.environ.get('HF_CONV_PAIR', '1'), ts[len(ts) // 2], tiles, nkb))
    if int(os.environ.get('HF_CONV_DBG', '0')) & 8:
        import ctypes
        buf = (ctypes.c_ulonglong * 16)()
        run()
        _lib.check(lib.hf_debug_conv_stamps(buf))
        t0 = buf[0]
        names_ = ['entry', 'setup done', 'dep wait done', 'first stage landed', 'tile0 MMAs issued', 'tile0 acc complete',
                  'tile0 converted', 'last store issued', 'last store landed', 'roles done']
        print('   ' + '  '.join('%s +%.2f' % (n, (buf[i] - t0) / 1e3) for i, n in enumerate(names_)))
