"""humaniflow_b200 — B200-native implementation of HuManiFlow's per-image sampling hot path.

Public surface (mirrors the reference's two classes):
    HumaniflowModel(device, model_cfg, smpl_parents)   <- models/humaniflow_model.py
    SMPL(model_path, batch_size, gender, num_betas)    <- models/smpl.py
    get_humaniflow_cfg_defaults()                      <- configs/humaniflow_config.py (MODEL subtree)
All compute is in libhumaniflow_b200.so (hand-written sm_100a CUDA) behind include/humaniflow_b200.h.
"""
from .config import get_humaniflow_cfg_defaults, get_model_cfg_defaults, Node  # noqa: F401
from .humaniflow_model import HumaniflowModel, immediate_parent_to_all_ancestors  # noqa: F401
from .smpl import SMPL, SMPLOutput  # noqa: F401

__all__ = ['HumaniflowModel', 'SMPL', 'SMPLOutput', 'get_humaniflow_cfg_defaults', 'get_model_cfg_defaults',
           'immediate_parent_to_all_ancestors', 'Node']
