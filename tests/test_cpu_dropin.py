"""The 'unchanged scripts' boundary: with ``humaniflow_b200/dropin`` ahead of the
reference checkout on PYTHONPATH, the import block of scripts/run_predict.py:7-12 / run_evaluate.py:6-16 must resolve
``models.humaniflow_model`` / ``models.smpl`` to this package and every other ``models.*`` module to the checkout.

Runs the import block in a subprocess the way the scripts are started (``python scripts/run_x.py`` from the checkout
root, so sys.path = [scripts/, PYTHONPATH..., '.']): once against a fake checkout laid out like the reference (regular
``models`` package with an empty ``__init__.py``), and once against /root/reference itself when it is present."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, 'humaniflow_b200', 'dropin')

IMPORT_BLOCK = textwrap.dedent('''
    import sys
    sys.path.append('.')                                               # scripts/run_predict.py:7
    from models.humaniflow_model import HumaniflowModel                # :9
    from models.smpl import SMPL                                       # :10
    from models.pose2D_hrnet import PoseHighResolutionNet              # :11
    from models.canny_edge_detector import CannyEdgeDetector           # :12
    import models
    print(HumaniflowModel.__module__, SMPL.__module__, PoseHighResolutionNet.__module__, CannyEdgeDetector.__module__)
    assert HumaniflowModel.__module__ == 'humaniflow_b200.humaniflow_model'
    assert SMPL.__module__ == 'humaniflow_b200.smpl'
    assert PoseHighResolutionNet.__module__ == 'models.pose2D_hrnet' and CannyEdgeDetector.__module__ == 'models.canny_edge_detector'
    import models.humaniflow_model as hm
    assert hasattr(hm, 'immediate_parent_to_all_ancestors')
''')


def _run(checkout, script_dir):
    os.makedirs(script_dir, exist_ok=True)
    script = os.path.join(script_dir, 'run_imports.py')
    with open(script, 'w') as f:
        f.write(IMPORT_BLOCK)
    env = dict(os.environ, PYTHONPATH=DROPIN + os.pathsep + ROOT)
    return subprocess.run([sys.executable, '-W', 'ignore', script], cwd=checkout, env=env, capture_output=True, text=True, timeout=300)


def test_import_block_against_fake_checkout(tmp_path):
    co = tmp_path / 'HuManiFlow'
    (co / 'models').mkdir(parents=True)
    (co / 'models' / '__init__.py').write_text('')                     # the reference's models/ is a regular package
    (co / 'models' / 'pose2D_hrnet.py').write_text('class PoseHighResolutionNet:\n    pass\n')
    (co / 'models' / 'canny_edge_detector.py').write_text('class CannyEdgeDetector:\n    pass\n')
    # the checkout's own implementations must be shadowed, not imported (they need pyro / smplx)
    (co / 'models' / 'humaniflow_model.py').write_text('raise ImportError("reference humaniflow_model imported")\n')
    (co / 'models' / 'smpl.py').write_text('raise ImportError("reference smpl imported")\n')
    r = _run(str(co), str(co / 'scripts'))
    assert r.returncode == 0, r.stderr[-2000:]


@pytest.mark.skipif(not os.path.isdir('/root/reference/models'), reason='reference checkout not present')
def test_import_block_against_the_reference_checkout(tmp_path):
    r = _run('/root/reference', str(tmp_path / 'scripts'))
    assert r.returncode == 0, r.stderr[-2000:]
