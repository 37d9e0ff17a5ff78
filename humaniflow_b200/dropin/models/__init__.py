"""Import shim: put ``humaniflow_b200/dropin`` ahead of the reference checkout on ``sys.path`` and the
reference's ``from models.humaniflow_model import HumaniflowModel`` / ``from models.smpl import SMPL``
(scripts/run_predict.py:9-10, run_evaluate.py:14-15) resolve to the B200 implementation.

The reference's ``models`` is a regular package (empty ``__init__.py``), so a package of the same name placed first on
``sys.path`` would hide all of it.  This one therefore appends every other ``models`` directory found on ``sys.path``
(the checkout the scripts add with ``sys.path.append('.')``) to its own ``__path__``: submodules that exist here
(``humaniflow_model``, ``smpl``) win, everything else (``canny_edge_detector``, ``pose2D_hrnet``, ``resnet``,
``norm_flows`` ...) keeps coming from the reference.  See INTEGRATION.md.
"""
import pkgutil

__path__ = pkgutil.extend_path(__path__, __name__)
