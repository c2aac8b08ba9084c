"""Drop-in ``networks`` package: same module/class names, constructors, ``state_dict`` layout and
``forward(ctx_x, ctx_y, tgt_x, test=False) -> (mu, None, 0)`` contract as the reference's
``networks/``, with every hot-path op running in libb200np.so (sm_100a CUDA).

Only the six hot-path modules live here (``ANPDistractor, ANP, ANPShapeNet1D, CNPDistractor,
CondNeuralProcess, CNPShapeNet1D`` + private helpers).  ``__path__`` is extended with every other
``networks`` directory on ``sys.path``, so with this package FIRST on the path and the reference
checkout behind it, ``networks.<anything else>`` (``models``, ``ResNet``, ``fast_attention``, the
MAML / MR / FCL model files) still resolves to the reference's own files, while the reference's
``train.py:41-45`` importlib lookup of a hot-path method resolves to the classes here.
``b200_run.py`` (next to this package) sets the path up that way for ``train.py`` /
``evaluation.py`` / ``refinement.py``.
"""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
