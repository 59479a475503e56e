"""ORACLE - test infrastructure, not product.

CPU restatement of the reference's view-acquisition path (multimodallearning/
acquisition-focus, pure Python over torch ATen).  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import anything from here, and only as the checker / the
timed CPU baseline - never as the thing shipped.  The product
(``acquisition_focus_b200``) does not import this package and fails loudly when
its CUDA library is missing.

Parity pinning: the reference has no tests or golden vectors of its own
(``/root/reference/tests/__init__.py`` is empty), so the oracle is pinned
against outputs of the reference itself run on CPU in the build container
(``oracle/make_golden.py`` -> ``tests/golden/*.npz``; bitwise-equal forward
results, see ``tests/test_oracle_golden.py``), and ``oracle/aten_np.py`` is
pinned bitwise against torch-CPU ATen.

Modules: ``ref_import`` (import the unmodified reference, build container
only), ``af_oracle`` (torch-CPU port that travels), ``aten_np`` (numpy bit-level
restatement of ATen affine_grid / grid_sampler_3d), ``cases`` (deterministic
inputs), ``make_golden`` (fixture generator).
"""
