"""Deterministic weights / inputs shared by tests/golden/make_golden.py and the tests.

numpy's legacy RandomState stream is stable across numpy versions, so fixtures only need to store
OUTPUTS; inputs and weights are regenerated from seeds.
"""
import numpy as np
import torch


def fill_state_dict(shapes, seed, scale_fn=None):
    """shapes: ordered dict name -> shape.  Returns name -> fp32 tensor.  BatchNorm statistics get
    non-trivial values (mean~N(0,.1), var~U(.5,1.5), weight~U(.5,1.5), bias~N(0,.1)) so BN is not a no-op
    (SURVEY 8d config 3); conv / linear weights ~ N(0, 1/fan_in) * gain."""
    rs = np.random.RandomState(seed)
    out = {}
    for name, shape in shapes.items():
        shape = tuple(shape)
        if name.endswith('num_batches_tracked'):
            out[name] = torch.zeros((), dtype=torch.long)
            continue
        if name.endswith('running_var'):
            a = rs.uniform(0.5, 1.5, size=shape)
        elif name.endswith('running_mean'):
            a = rs.standard_normal(shape) * 0.1
        elif '.bn' in name or 'downsample.1' in name or name.startswith('bn') or '.bn' in ('.' + name):
            a = rs.uniform(0.5, 1.5, size=shape) if name.endswith('weight') else rs.standard_normal(shape) * 0.1
        elif len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            a = rs.standard_normal(shape) * np.sqrt(2.0 / fan_in)
        else:
            a = rs.standard_normal(shape) * 0.05
        out[name] = torch.tensor(np.asarray(a, dtype=np.float32).reshape(shape))
    return out


def det_input(shape, seed, kind='normal'):
    rs = np.random.RandomState(seed)
    a = rs.standard_normal(shape) if kind == 'normal' else rs.uniform(0, 1, size=shape)
    return torch.tensor(a.astype(np.float32))
