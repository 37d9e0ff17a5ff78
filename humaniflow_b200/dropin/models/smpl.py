from humaniflow_b200.smpl import SMPL, SMPLOutput  # noqa: F401
