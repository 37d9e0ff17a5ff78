"""Synthetic SMPL-shaped model data and inputs.

The real SMPL_{NEUTRAL,MALE,FEMALE}.pkl files are licence-gated and absent, so tests and
bench.py run on SMPL-*shaped* buffers generated here: 6890 vertices, 24 joints,
10 betas, 207 pose-feature dims, <=4 skinning weights per vertex, standard SMPL kinematic tree.
The generator is committed instead of the arrays.  numpy RandomState keeps the stream stable.
"""
import numpy as np
import torch

SMPL_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]
NUM_VERTS = 6890


def _sparse_rows(rs, rows, cols, nnz_per_row):
    a = np.zeros((rows, cols), dtype=np.float64)
    for r in range(rows):
        idx = rs.choice(cols, size=nnz_per_row, replace=False)
        w = rs.uniform(0.1, 1.0, size=nnz_per_row)
        a[r, idx] = w / w.sum()
    return a


def synthetic_smpl_data(seed=0, num_verts=NUM_VERTS, num_betas=10, regressors=None):
    """Returns a dict of fp32 torch tensors with the smplx buffer names/shapes
    (v_template, shapedirs, posedirs, J_regressor, lbs_weights, parents) plus the three extra regressors
    of models/smpl.py:16-25 (synthetic sparse rows unless ``regressors`` supplies the real ones)."""
    rs = np.random.RandomState(seed)
    V = num_verts
    d = {}
    d['v_template'] = rs.standard_normal((V, 3)) * 0.3
    d['shapedirs'] = rs.standard_normal((V, 3, num_betas)) * 0.01
    d['posedirs'] = rs.standard_normal((207, V * 3)) * 0.003
    d['J_regressor'] = _sparse_rows(rs, 24, V, 32)
    W = np.zeros((V, 24))
    for v in range(V):
        k = rs.randint(1, 5)
        idx = rs.choice(24, size=k, replace=False)
        w = rs.uniform(0.05, 1.0, size=k)
        W[v, idx] = w / w.sum()
    d['lbs_weights'] = W
    d['J_regressor_extra'] = _sparse_rows(rs, 9, V, 7)
    d['J_regressor_cocoplus'] = _sparse_rows(rs, 19, V, 5)
    d['J_regressor_h36m'] = _sparse_rows(rs, 17, V, 6)
    if regressors is not None:
        d.update(regressors)
    out = {k: torch.tensor(np.asarray(v, dtype=np.float32)) for k, v in d.items()}
    out['parents'] = list(SMPL_PARENTS)
    return out


def synthetic_proxy_input(batch, channels=18, size=256, seed=0):
    """(B,C,size,size) fp32 in [0,1): sparse non-negative 'edge' channel 0, smooth blobs elsewhere
    (SURVEY 8d config 3).  Cheap to generate; values only matter for being finite and non-constant."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(batch, channels, size, size, generator=g)
    x[:, 0] = (x[:, 0] > 0.9).float() * x[:, 0]
    return x
