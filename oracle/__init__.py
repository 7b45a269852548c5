"""CPU oracle for the SMC hot path of nchopin/particles -- TEST INFRASTRUCTURE ONLY.

This package is a restatement (numpy + a few lines of plain C) of the per-step
algorithm of ``particles.core.SMC`` and of the numerics underneath it
(``particles/resampling.py``, ``particles/distributions.py``,
``particles/state_space_models.py``, ``particles/kalman.py``); every function
cites the reference file:line it follows.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the live
reference (``PYTHONPATH=/root/reference``) in the build container, runs it with
fixed ``numpy.random`` seeds and stores its outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks that this oracle reproduces those outputs
bit-for-bit (same legacy MT19937 stream, consumed in the same order).

Who may import this package: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- as the checker
or as the timed CPU baseline, never as the product.  Nothing under
``particles_b200/`` imports it; the product path raises if the CUDA library is
missing rather than falling back to this code.
"""
