from humaniflow_b200.humaniflow_model import HumaniflowModel, immediate_parent_to_all_ancestors  # noqa: F401
