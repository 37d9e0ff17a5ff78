"""Import shim: put ``humaniflow_b200/dropin`` ahead of the reference checkout on ``sys.path`` and the
reference's ``from models.humaniflow_model import HumaniflowModel`` / ``from models.smpl import SMPL``
(scripts/run_predict.py:9-10, run_evaluate.py:14-15) resolve to the B200 implementation.  See INTEGRATION.md."""
