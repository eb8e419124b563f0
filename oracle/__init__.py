"""CPU oracle for the PINN training hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this package.  The product path (``pinns-tf2.0_b200/``) never does and fails
loudly when its CUDA library is missing.

PARITY UNPINNED: the reference (pierremtb/PINNs-TF2.0) ships no tests, golden vectors or fixtures for
this path, and its arithmetic lives in ``tensorflow==2.0.0-rc0`` which is not installable here
(SURVEY.md section 8(c)).  The oracle is therefore pinned only by (i) line-by-line restatement of the
reference Python with citations, and (ii) agreement between two independent formulations:
``oracle.reference_port`` (nested reverse-mode autograd, same structure as the nested GradientTapes) and
``oracle.taylor`` (closed-form forward Taylor-mode + hand-derived reverse sweep, numpy only).
"""
