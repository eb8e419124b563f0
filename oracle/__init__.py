"""CPU oracle for the PINN training hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this package.  The product path (``pinns-tf2.0_b200/``) never does and fails
loudly when its CUDA library is missing.

PARITY STATUS.  The reference (pierremtb/PINNs-TF2.0) ships no tests, golden vectors or fixtures for this path, and its
arithmetic lives in ``tensorflow==2.0.0-rc0``, which is not installable here: against TensorFlow's
own kernels parity is UNPINNED and stays so.  What the oracle IS pinned to:

(i)   the reference's own Python, executed: ``tests/golden/make_reference_fixtures.py`` imports the unmodified
      ``utils/neuralnetwork.py``, ``utils/custom_lbfgs.py``, ``utils/logger.py`` and the ``*InformedNN`` classes of the example
      scripts from /root/reference and runs ``grad()``, ``fit()``, ``lbfgs()``, ``predict()`` on the golden inputs, with
      ``import tensorflow`` resolved to the API emulation in ``oracle/tf_emulation`` (tape/Keras/Adam semantics on torch fp64).
      The results are committed as ``tests/golden/reference_run.npz`` and equal this oracle's golden values to <= 1e-16
      (``tests/test_reference_pin.py``);
(ii)  line-by-line restatement of the reference Python with file:line citations (``oracle.reference_port``: nested
      reverse-mode autograd, same structure as the nested GradientTapes);
(iii) agreement with an independent formulation (``oracle.taylor``: closed-form forward Taylor-mode + hand-derived reverse
      sweep, numpy only).
"""
