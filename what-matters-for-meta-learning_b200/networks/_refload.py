"""Load the reference's own copy of a ``networks`` module that this package shadows (found behind this package on
``networks.__path__`` when the reference checkout is on ``sys.path``).  ``name`` may be dotted (``"bbb.BBBConv"``)."""
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def reference_module(name):
    import networks
    parts = name.split(".")
    for d in list(networks.__path__):
        cand = os.path.join(d, *parts) + ".py"
        if os.path.abspath(d) != _HERE and os.path.isfile(cand):
            package = ".".join(["networks"] + parts[:-1])
            full = f"{package}._reference_{parts[-1]}"
            if full in sys.modules:
                return sys.modules[full]
            spec = importlib.util.spec_from_file_location(full, cand)
            mod = importlib.util.module_from_spec(spec)
            mod.__package__ = package              # its relative imports (`from .maml_model import Model`) keep working
            sys.modules[full] = mod
            spec.loader.exec_module(mod)
            return mod
    return None
