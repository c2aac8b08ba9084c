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
through the committed golden vectors that ``tests/golden/make_golden.py`` (2-task cases) and
``tests/golden/make_golden_full.py`` (BASELINE.json's full sizes) produced by running the unmodified
reference: ``tests/test_oracle_golden.py`` and ``tests/test_parity_at_size.py``.
``make_ref``    copies the unmodified reference into the git-ignored ``oracle/_ref/`` so that it travels
                to the GPU box: CPU / on-GPU reference arms of ``bench.py`` and the unmodified
                ``ModelTrainer`` that ``tests/test_dropin.py`` drives on the drop-in modules.
"""
