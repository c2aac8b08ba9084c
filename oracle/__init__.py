"""Test oracle for the task-batched neural-process hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / the timed CPU baseline.  Nothing under
``what-matters-for-meta-learning_b200/`` imports it; the product path raises when the CUDA
library is missing instead of falling back to this code.

Contents
--------
``np_oracle``   CPU (torch fp32/fp64) restatement of the reference algorithm, every function
                citing the reference file:line it follows.
``synth``       integer-hash synthetic task batches (bit-reproducible on any host, no RNG).
``ref_shims``   import stubs that let the *real* reference under /root/reference be imported in
                the build container; used to pin ``np_oracle`` and to generate
                ``tests/golden/*.npz`` (script: ``tests/golden/make_golden.py``).  The reference
                does not exist on the GPU box, so nothing run there touches ``ref_shims``.

Parity status: **pinned** -- ``np_oracle`` is checked against the live reference modules
(tests/test_oracle_vs_reference.py, runs only where /root/reference exists) and against the
committed golden vectors produced by the live reference (tests/test_oracle_golden.py).
"""
