#!/usr/bin/env python
"""Run one of the reference's entry scripts on the B200 path:

    python /path/to/what-matters-for-meta-learning_b200/b200_run.py [--reference-root DIR] train.py --config cfg/...yaml

``python train.py`` alone cannot pick the drop-in up: Python puts the script's directory ahead of
``PYTHONPATH`` and the reference's ``networks`` is a regular package, so ``networks.<method>``
(``train.py:41-45``) would resolve to the reference's classes.  This launcher puts the drop-in
package directory FIRST on ``sys.path`` and the reference checkout right behind it, then runs the
script as ``__main__``.  The drop-in ``networks`` / ``trainer`` packages extend their ``__path__`` with
the reference's directories, so everything that is not on the hot path (``trainer.model_trainer``,
``networks.MAMLShapeNet1D``, ``dataset``, ``configs``, ``utils``, ``evaluator``) is the reference's own code.
"""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def setup_path(reference_root):
    """sys.path = [drop-in package, reference checkout, ...rest]; returns the reference root used."""
    reference_root = os.path.abspath(reference_root)
    for p in (HERE, reference_root):
        while p in sys.path:
            sys.path.remove(p)
    sys.path[:0] = [HERE, reference_root]
    for name in ("networks", "trainer"):   # a copy imported before the path was arranged would stay cached
        for k in [k for k in sys.modules if k == name or k.startswith(name + ".")]:
            del sys.modules[k]
    return reference_root


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    ref = os.getcwd()
    if argv and argv[0] == "--reference-root":
        ref = argv[1]
        argv = argv[2:]
    if not argv:
        raise SystemExit(__doc__)
    script = argv[0] if os.path.isabs(argv[0]) else os.path.join(ref, argv[0])
    if not os.path.isfile(script):
        raise SystemExit(f"b200_run: {script} not found (run from the reference checkout or pass --reference-root)")
    setup_path(ref)
    # forward / backward as CUDA graphs cached per (nc, nt) shape behind the unchanged nn.Module contract
    # (b200np/graphed.py); B200NP_GRAPHS=0 keeps the eager launches
    os.environ.setdefault("B200NP_GRAPHS", "1")
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")   # a plain script file: run_path leaves sys.path alone


if __name__ == "__main__":
    main()
