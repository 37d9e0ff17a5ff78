"""CPU oracle for the HuManiFlow per-image sampling hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / the timed CPU baseline.  Nothing under
``humaniflow_b200/`` imports it; the product path fails loudly when the CUDA library is missing.

What it is: a plain PyTorch-on-CPU restatement (fp32, fp64 where the reference uses fp64) of the
path ``encoder -> heads -> ancestor-conditioned SO(3) flow (sample / log_prob) -> SMPL LBS``.

Parity status (be precise about what is pinned):

* PINNED against the real reference, run in the build container by
  ``tests/golden/make_golden.py`` (vectors committed under ``tests/golden/``):
  ``oracle.so3`` (rot6d, so3_exp/log/log_pi/xset/log|detJ|  <- utils/rigid_transform_utils.py),
  ``oracle.flow.radial_tanh_*`` (<- models/norm_flows/transforms/scaled_radial_tanh_transform.py),
  ``oracle.resnet`` (<- models/resnet.py), and the pre-image/logsumexp plumbing of
  ``oracle.flow.so3_flow_log_prob`` (<- models/norm_flows/local_diffeo_transformed_distribution.py,
  so3_exp_transform.py, to_transform.py; the reference classes are imported with name-only stubs for the
  two pyro base classes they subclass -- the stubs contain no arithmetic).
* PARITY UNPINNED for the arithmetic that lives in un-vendored third-party packages that are absent
  from /root/reference and not installable here (no network):
  pyro-ppl==1.7.0 (``SplineCoupling`` / ``ConditionalSpline`` / ``_monotonic_rational_spline``,
  ``ConditionalDenseNN``, ``Permute``) and smplx==0.1.26 (``lbs``, ``batch_rodrigues``,
  ``batch_rigid_transform``, ``vertices2joints``, ``VertexJointSelector``, ``SMPL.forward``).
  ``oracle.spline`` and ``oracle.smpl`` restate their published algorithms; the reference has no tests
  or golden vectors for them, so these are checked only by oracle-free invariants
  (inverse(forward(x)) == x, log-dets cancel, density integrates, LBS at identity == v_shaped, ...).
* SMPL model files are licence-gated and absent: all LBS data are synthetic SMPL-shaped buffers
  (``humaniflow_b200.synthetic``) plus the three real regressors shipped with the reference
  (copied numerically into tests/golden as sparse triplets by make_golden.py).
"""
