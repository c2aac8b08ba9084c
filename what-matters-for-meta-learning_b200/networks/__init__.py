"""Drop-in ``networks`` package: same module/class names, constructors, ``state_dict`` layout and
``forward(ctx_x, ctx_y, tgt_x, test=False) -> (mu, None, 0)`` contract as the reference's
``networks/``, with every hot-path op running in libb200np.so (sm_100a CUDA).

Put this directory first on ``sys.path`` and the reference's ``train.py:41-45`` importlib lookup
(``networks.<method>`` -> class ``<method>``) resolves to these classes.
"""
