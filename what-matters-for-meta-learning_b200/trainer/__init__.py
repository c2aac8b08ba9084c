"""Drop-in ``trainer`` package: only ``trainer.losses`` (LossFunc, reference ``train.py:92`` /
``model_trainer.py:77``) is replaced.  ``__path__`` is extended with the reference's ``trainer``
directory (a namespace package there: it has no ``__init__.py``) when the reference checkout is on
``sys.path``, so ``trainer.model_trainer``, ``trainer.maml_trainer``, ``trainer.meta_learner_reg``,
``trainer.mmaml_trainer`` and ``trainer.base_trainer`` (``train.py:23-30``) keep resolving to the
reference's own, unmodified files.
"""
import os
import sys

from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
# pkgutil.extend_path only adds directories that contain an __init__.py; the reference's ``trainer`` has none
for _d in sys.path:
    _cand = os.path.join(_d or ".", "trainer")
    if os.path.isfile(os.path.join(_cand, "model_trainer.py")) and _cand not in __path__:
        __path__.append(_cand)
