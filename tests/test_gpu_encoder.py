"""ResNet encoder on tcgen05/TMA: single fused convolutions and the whole trunk vs CPU references.

Numerics contract of the CUDA encoder: bf16 operands, fp32 accumulation, bf16 activations between
layers.  Tolerances: a single conv vs the same contract on the CPU: 1 bf16 ulp (relative 2^-7) + 1e-3 abs;
whole trunk vs the contract restated on the CPU (util.resnet_bf16emu): rel L2 <= 7e-3;
whole trunk vs the plain fp32 oracle (oracle.resnet, pinned to the reference class): rel L2 <= 3e-2.
"""
import ctypes

import pytest
import torch

from humaniflow_b200 import _lib
from oracle.resnet import resnet_forward
from util import bf16r, conv_bf16_ref, make_model, resnet_bf16emu

pytestmark = pytest.mark.gpu


def _bits(t):
    return t.to(torch.bfloat16).contiguous().view(torch.int16)


def _run_conv(x, w, bias, res, stride, pad, relu, impl):
    lib = _lib.load()
    B, H, W, Ci = x.shape
    Co, k = w.shape[0], w.shape[1]
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    xd, wd = _bits(x).cuda(), _bits(w).cuda()
    bd = bias.float().cuda()
    rd = None if res is None else _bits(res).cuda()
    y = torch.empty(B, Ho, Wo, Co, dtype=torch.int16, device='cuda')
    _lib.check(lib.hf_conv2d_nhwc(_lib.ptr(xd), _lib.ptr(wd), _lib.ptr(bd), _lib.ptr(rd), _lib.ptr(y), B, H, W, Ci, Co, k,
                                  stride, pad, int(relu), impl, _lib.stream()))
    torch.cuda.synchronize()
    return y.view(torch.bfloat16).float().cpu()


CASES = [
    # B, H, W, Cin, Cout, k, stride, pad, res, relu
    (2, 16, 16, 64, 64, 1, 1, 0, False, True),       # layer1 1x1
    (2, 16, 16, 64, 256, 1, 1, 0, True, True),       # 1x1 expand + residual
    (2, 16, 16, 64, 64, 3, 1, 1, False, True),       # 3x3, zero padding from TMA OOB fill
    (1, 32, 32, 128, 128, 3, 2, 1, False, True),     # 3x3 stride 2 (parity tensor maps)
    (2, 16, 16, 256, 512, 1, 2, 0, False, False),    # 1x1 stride 2 downsample, no relu
    (3, 8, 8, 512, 128, 3, 1, 1, True, True),        # 8x8 maps: two images per tile, odd batch -> partial tile
    (5, 4, 4, 128, 64, 3, 1, 1, False, False),       # 4x4 maps: 8 images per tile, ragged
    (1, 64, 64, 64, 128, 3, 1, 1, False, True),      # many tiles, BN=128
    (1, 14, 14, 64, 64, 3, 1, 1, False, True),       # non power-of-two maps: partial tiles in w and h
    (8, 64, 64, 64, 256, 1, 1, 0, False, True),      # >= 148 wide tiles: the BN=256 kernel variant, persistent loop > 1 tile/CTA
    (5, 64, 64, 64, 256, 1, 1, 0, True, True),       # residual ring (3 staging tiles) over many tiles per CTA
    (6, 32, 32, 128, 512, 3, 2, 1, False, False),    # BN=256 with a K-heavy strided 3x3
]


@pytest.mark.parametrize('case', CASES)
def test_single_conv_tcgen05(case):
    B, H, W, Ci, Co, k, s, p, with_res, relu = case
    g = torch.Generator().manual_seed(hash(case) % 1000)
    x = bf16r(torch.randn(B, H, W, Ci, generator=g))
    w = bf16r(torch.randn(Co, k, k, Ci, generator=g) / (k * k * Ci) ** 0.5)
    bias = torch.randn(Co, generator=g) * 0.1
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    res = bf16r(torch.randn(B, Ho, Wo, Co, generator=g)) if with_res else None
    ref = conv_bf16_ref(x, w, bias, res, s, p, relu)
    got = _run_conv(x, w, bias, res, s, p, relu, impl=0)
    assert got.shape == ref.shape
    err = (got - ref).abs()
    assert (err <= 2 ** -7 * ref.abs() + 1e-3).all(), (err.max().item(), case)
    simt = _run_conv(x, w, bias, res, s, p, relu, impl=1)
    assert ((simt - ref).abs() <= 2 ** -7 * ref.abs() + 1e-3).all()


@pytest.mark.parametrize('layers,size,B', [(18, 64, 2), (50, 64, 3), (50, 256, 2)])
def test_trunk(layers, size, B):
    m, sd, cfg = make_model(layers, seed=20 + layers)
    enc = m.image_encoder.cuda()
    g = torch.Generator().manual_seed(5)
    x = torch.rand(B, 18, size, size, generator=g)
    with torch.no_grad():
        f32 = resnet_forward(sd, x, layers)
        emu = resnet_bf16emu(sd, x, layers)
    got = enc(x.cuda()).cpu()
    print('stem channels per pixel:', _lib.load().hf_encoder_stem_channels(enc._enc))
    assert got.shape == f32.shape and got.dtype == torch.float32
    rel = lambda a, b: ((a - b).norm() / b.norm()).item()
    # two valid bf16 implementations differ by their fp32 summation order inside every k-reduction (here: even / odd k-blocks
    # accumulate separately), which re-rounds a few activations per layer; after 53 layers that is ~5e-3 rel-L2
    assert rel(got, emu) <= 7e-3, rel(got, emu)
    assert rel(got, f32) <= 3e-2, rel(got, f32)
    # cross-check the tensor-core path against the SIMT direct convolution on the same packed weights
    enc.set_impl(1)
    simt = enc(x.cuda()).cpu()
    enc.set_impl(0)
    assert rel(got, simt) <= 7e-3, rel(got, simt)


def test_model_with_image_input():
    """Full predict path (BASELINE configs[2]) at a small size: image -> encoder -> flow; return-dict contract."""
    m, sd, cfg = make_model(50, seed=30)
    m = m.cuda()
    x = torch.rand(2, 18, 256, 256, generator=torch.Generator().manual_seed(1))
    out = m(x.cuda(), num_samples=5, return_input_feats=True)
    assert out['input_feats'].shape == (2, 2048)
    assert out['pose_rotmats_samples'].shape == (2, 5, 23, 3, 3)
    assert out['cam_wp'].shape == (2, 3) and out['glob_rotmat'].shape == (2, 3, 3)
    assert out['shape_mode'].shape == (2, 10) and out['shape_log_std'].shape == (2, 10)
    assert hasattr(out['shape_dist_for_loglik'], 'log_prob')
    for k, v in out.items():
        if torch.is_tensor(v):
            assert torch.isfinite(v).all(), k
    with pytest.raises(RuntimeError):
        m.train()
        m(x.cuda())


BRANCH_CASES = [
    # B, H, W, Cin, Cout, k, stride (main, always 1), Cin2, stride2 (branch), H2 = H * stride2
    (2, 16, 16, 64, 256, 1, 64, 1),        # layer1 block 0: conv3 (64 -> 256) + 1x1 downsample of the block input (64 -> 256)
    (2, 16, 16, 128, 512, 1, 256, 2),      # layer2 block 0: conv3 + stride-2 downsample of the 32x32 block input
    (3, 8, 8, 256, 1024, 1, 512, 2),       # layer3 block 0, odd batch: partial tile, BN = 256
    (2, 8, 8, 128, 128, 3, 64, 2),         # BasicBlock (ResNet-18): 3x3 conv2 + stride-2 1x1 downsample
]


@pytest.mark.parametrize('case', BRANCH_CASES)
def test_fused_branch_conv(case):
    """hf_enc_op.src2: the block's 1x1 downsample branch accumulated into the last conv's fp32 accumulator
    (models/resnet.py:108-118: out = bn3(conv3(out)) + downsample(x); relu).  Reference = both convolutions in fp32 on the
    bf16-rounded operands, summed BEFORE the single rounding to bf16."""
    import torch.nn.functional as F
    B, H, W, Ci, Co, k, Ci2, s2 = case
    lib = _lib.load()
    g = torch.Generator().manual_seed(sum(case))
    x = bf16r(torch.randn(B, H, W, Ci, generator=g))
    x2 = bf16r(torch.randn(B, H * s2, W * s2, Ci2, generator=g))
    w1 = bf16r(torch.randn(Co, k, k, Ci, generator=g) / (k * k * Ci) ** 0.5)
    w2 = bf16r(torch.randn(Co, 1, 1, Ci2, generator=g) / Ci2 ** 0.5)
    bias = torch.randn(Co, generator=g) * 0.1
    y1 = F.conv2d(x.permute(0, 3, 1, 2), w1.permute(0, 3, 1, 2), padding=k // 2)
    y2 = F.conv2d(x2.permute(0, 3, 1, 2), w2.permute(0, 3, 1, 2), stride=s2)
    ref = bf16r(F.relu((y1 + y2).permute(0, 2, 3, 1) + bias))
    wcat = torch.cat([w1.reshape(Co, -1), w2.reshape(Co, -1)], 1).contiguous()
    for impl in (0, 1):
        xd, x2d, wd, bd = _bits(x).cuda(), _bits(x2).cuda(), _bits(wcat).cuda(), bias.cuda()
        y = torch.empty(B, H, W, Co, dtype=torch.int16, device='cuda')
        _lib.check(lib.hf_conv2d_nhwc_branch(_lib.ptr(xd), _lib.ptr(wd), _lib.ptr(bd), None, _lib.ptr(y), B, H, W, Ci, Co, k, 1, k // 2, 1,
                                             _lib.ptr(x2d), H * s2, W * s2, Ci2, s2, impl, _lib.stream()))
        torch.cuda.synchronize()
        got = y.view(torch.bfloat16).float().cpu()
        err = (got - ref).abs()
        assert (err <= 2 ** -7 * ref.abs() + 1e-3).all(), (impl, err.max().item(), case)


@pytest.mark.parametrize('layers,size,B', [(18, 64, 3), (50, 256, 2)])
def test_staged_and_bf16_inputs_match_the_fp32_input_bit_for_bit(layers, size, B):
    """hf_proxy_rep_staged writes the proxy representation straight into the encoder's bf16 NHWC stem input
    (no fp32 NCHW intermediate, no layout pass); the bf16 NCHW entry point halves the host->device bytes.  Both round the
    same fp32 values to bf16 with the same rounding, so the features must be IDENTICAL to the fp32 NCHW path."""
    from humaniflow_b200.proxy_rep import build_proxy_representation
    from humaniflow_b200.resnet import StagedInput
    m, sd, cfg = make_model(layers, seed=60 + layers, res_gain=0.3)
    enc = m.image_encoder.cuda()
    g = torch.Generator().manual_seed(7)
    rgb = torch.rand(B, 3, size, size, generator=g).cuda()
    j2d = (torch.rand(B, 17, 2, generator=g) * size).cuda()
    vis = (torch.rand(B, 17, generator=g) > 0.2).cuda()
    proxy = build_proxy_representation(rgb, j2d, vis)                       # (B,18,H,W) fp32
    ref = enc(proxy).clone()
    staged = build_proxy_representation(rgb, j2d, vis, encoder=enc)
    assert isinstance(staged, StagedInput) and staged.shape == (B, 18, size, size)
    got = enc(staged).clone()
    assert torch.equal(got, ref)
    again = enc(build_proxy_representation(rgb, j2d, vis, encoder=enc))   # the zero border survives repeated use
    assert torch.equal(again, ref)
    assert torch.equal(enc(proxy.to(torch.bfloat16)), ref)
    assert torch.equal(enc(proxy), ref)                                     # and the fp32 path after the staged one
    out = m.cuda()(staged, num_samples=2)
    assert out['pose_rotmats_samples'].shape == (B, 2, 23, 3, 3)
