"""Model configuration node with the fields HumaniflowModel reads.

Mirrors the MODEL subtree of /root/reference/configs/humaniflow_config.py:8-22.  The reference uses a yacs
``CfgNode``; anything with attribute access works here (a yacs node from the reference's own config
module, or this dependency-free ``Node``).
"""
import math


class Node(dict):
    """Minimal attribute-access dict (yacs is not a dependency of the hot path)."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__

    def clone(self):
        return Node({k: (v.clone() if isinstance(v, Node) else (list(v) if isinstance(v, list) else v))
                     for k, v in self.items()})


def get_model_cfg_defaults():
    """Reference defaults (ResNet-18; BASELINE.json's benchmark uses NUM_RESNET_LAYERS=50)."""
    return Node(
        NUM_IN_CHANNELS=18,
        NUM_RESNET_LAYERS=18,
        INPUT_SHAPE_GLOB_CAM_FEATS_DIM=256,
        NUM_SMPL_BETAS=10,
        NORM_FLOW=Node(
            CONTEXT_DIM=64,
            NUM_TRANSFORMS=2,
            TRANSFORM_TYPE='spline_coupling',
            TRANSFORM_NN_HIDDEN_DIMS=[64, 32, 32],
            NUM_SPLINE_SEGMENTS=8,
            PERMUTE_TYPE='permute',
            PERMUTE_NN_HIDDEN_DIMS=None,
            COMPACT_SUPPORT_RADIUS=1.5 * math.pi,
            BASE_DIST_STD=0.6,
        ),
    )


def get_humaniflow_cfg_defaults():
    """Same entry-point name as configs/humaniflow_config.py:109; only the MODEL subtree is populated."""
    return Node(MODEL=get_model_cfg_defaults())
