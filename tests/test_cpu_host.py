"""CPU-side checks: the C-ABI library loads and exports every declared symbol, the host mirrors keep the
reference's interface (state-dict keys, ancestors, config fields), and the product refuses to run without CUDA."""
import os
import re

import pytest
import torch

import humaniflow_b200 as hb
from humaniflow_b200 import _lib
from humaniflow_b200.synthetic import SMPL_PARENTS, synthetic_smpl_data

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from humaniflow_b200.build import build
    build()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, 'include', 'humaniflow_b200.h')).read()
    declared = set(re.findall(r'\b(hf_[a-z0-9_]+)\s*\(', header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.hf_version() == 100
    assert lib.hf_launch_count() == 0


def test_no_oracle_import_in_product():
    for root, _, files in os.walk(os.path.join(ROOT, 'humaniflow_b200')):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(root, f)).read()
                assert 'oracle' not in src.replace('the CPU oracle', ''), f


def test_state_dict_layout_matches_reference():
    """546 keys / 12,490,497 params (ResNet-18) and 744 / 27,065,601 (ResNet-50), observed by instantiating the
    reference class."""
    cfg = hb.get_model_cfg_defaults()
    for layers, nkeys, nparams in ((18, 546, 12490497), (50, 744, 27065601)):
        cfg.NUM_RESNET_LAYERS = layers
        m = hb.HumaniflowModel('cpu', cfg, SMPL_PARENTS)
        sd = m.state_dict()
        assert len(sd) == nkeys and sum(p.numel() for p in m.parameters()) == nparams
        for k in ('init_glob', 'init_cam', 'fc1.weight', 'fc_shape.bias', 'fc_glob.weight', 'fc_cam.weight',
                  'fc_input_shape_glob_cam_feats.weight', 'fc_flow_context.22.bias',
                  'pose_so3flow_transform_modules.45.nn.layers.3.weight', 'image_encoder.conv1.weight',
                  'image_encoder.layer4.0.downsample.1.running_var', 'image_encoder.bn1.num_batches_tracked'):
            assert k in sd, k
        assert sd['init_glob'].tolist() == [[1.0, 0.0, 0.0, 1.0, 0.0, 0.0]] and torch.allclose(sd['init_cam'], torch.tensor([0.9, 0.0, 0.0]))
        assert sd['pose_so3flow_transform_modules.0.nn.layers.0.weight'].shape == (64, 65)
        assert sd['pose_so3flow_transform_modules.0.nn.layers.3.weight'].shape == (62, 32)
        assert sd['fc_flow_context.22.weight'].shape == (64, 256 + 9 * 7)
        assert hasattr(m.pose_so3flow_transform_modules, 'eval') and all(hasattr(d, 'clear_cache') for d in m.pose_SO3flow_dists)


def test_ancestors():
    anc = hb.immediate_parent_to_all_ancestors(SMPL_PARENTS)
    assert [len(anc[j]) for j in range(23)] == [0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 3, 3, 3, 4, 4, 4, 5, 5, 6, 6, 7, 7]
    assert anc[22] == [20, 18, 16, 13, 8, 5, 2]
    from oracle.model import ancestors_of
    assert ancestors_of(SMPL_PARENTS) == [anc[j] for j in range(23)]


def test_refuses_cpu():
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    cfg = hb.get_model_cfg_defaults()
    m = hb.HumaniflowModel('cpu', cfg, SMPL_PARENTS).eval()
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        m(torch.zeros(1, 18, 64, 64))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        m(None, input_feats=torch.zeros(1, 512))
    smpl = hb.SMPL.from_arrays(synthetic_smpl_data(num_verts=256))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        smpl()


def test_unsupported_flow_variants_fail_loudly():
    cfg = hb.get_model_cfg_defaults()
    cfg.NORM_FLOW.TRANSFORM_TYPE = 'affine_coupling'
    with pytest.raises(NotImplementedError):
        hb.HumaniflowModel('cpu', cfg, SMPL_PARENTS)


def test_c_structs_match_the_header():
    """The ctypes mirrors of the C-ABI structs have the field count / order the header declares (hf_enc_op grew a fused
    second input: src2, cin2, stride2)."""
    import ctypes
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'include', 'humaniflow_b200.h')).read()
    body = re.search(r'typedef struct hf_enc_op \{(.*?)\} hf_enc_op;', hdr, re.S).group(1)
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    names = [n.strip() for decl in re.findall(r'int ([^;]+);', body) for n in decl.split(',')]
    assert names == [f[0] for f in _lib.EncOp._fields_], (names, _lib.EncOp._fields_)
    assert ctypes.sizeof(_lib.EncOp) == 4 * len(names)
    op = _lib.EncOp(_lib.OP_CONV, 0, 1, -1, 64, 64, 3, 1, 1, 1, 0)
    assert (op.src2, op.cin2, op.stride2) == (-1, 0, 1)          # defaults = no fused branch
