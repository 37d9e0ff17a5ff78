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


def _body_part_weights(rs, V):
    """Skinning weights with the structure of the real SMPL model: vertex ids fall into contiguous runs that belong to
    one body part (a primary joint); inside a run, patches of neighbouring vertices are bound to the same 1..4 joints
    taken from the primary joint and its kinematic neighbours (parent, grandparent, children), with per-vertex
    weights.  (The real SMPL_*.pkl are licence-gated and absent; their vertex numbering groups body parts in runs.)"""
    children = {j: [c for c, p in enumerate(SMPL_PARENTS) if p == j] for j in range(24)}
    W = np.zeros((V, 24))
    v = 0
    while v < V:
        run = min(V - v, int(rs.randint(60, 500)))
        prim = int(rs.randint(0, 24))
        nbrs = [SMPL_PARENTS[prim]] + children[prim] + ([SMPL_PARENTS[SMPL_PARENTS[prim]]] if SMPL_PARENTS[prim] > 0 else [])
        nbrs = [j for j in nbrs if j >= 0]
        end = v + run
        while v < end:
            patch = min(end - v, int(rs.randint(16, 120)))
            k = int(rs.randint(1, 5))
            others = list(rs.permutation(nbrs))[:k - 1]
            idx = [prim] + [int(j) for j in others]
            for u in range(v, v + patch):
                w = rs.uniform(0.05, 1.0, size=len(idx))
                W[u, idx] = w / w.sum()
            v += patch
    return W


def synthetic_smpl_data(seed=0, num_verts=NUM_VERTS, num_betas=10, regressors=None, skinning='random'):
    """Returns a dict of fp32 torch tensors with the smplx buffer names/shapes
    (v_template, shapedirs, posedirs, J_regressor, lbs_weights, parents) plus the three extra regressors
    of models/smpl.py:16-25 (synthetic sparse rows unless ``regressors`` supplies the real ones).
    ``skinning``: 'random' = every vertex bound to 1..4 joints drawn uniformly from the 24 (adversarial: no two
    neighbouring vertices share a joint set); 'body_parts' = SMPL-like structure (see _body_part_weights)."""
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
    d['lbs_weights'] = W if skinning == 'random' else _body_part_weights(np.random.RandomState(seed + 1000), V)
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
