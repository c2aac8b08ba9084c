"""Drop-in ``networks.bbb``: Bayes-by-backprop layers of the "MR" variants.  ``B200NP_BBB=1`` selects
the B200 layers (fused weight sampling + KL kernel in front of the library's convolution / GEMM kernels); otherwise the
package hands out the reference's own classes (found behind this package on ``networks.__path__``)."""
import os
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
import networks as _nw  # noqa: E402

for _d in list(_nw.__path__):          # the reference's networks/bbb directory, if its checkout is on the path
    _cand = os.path.join(_d, "bbb")
    if os.path.isdir(_cand) and _cand not in __path__:
        __path__.append(_cand)

from .misc import FlattenLayer, ModuleWrapper  # noqa: E402,F401
from .BBBConv import BBBConv2d  # noqa: E402,F401
from .BBBLinear import BBBLinear  # noqa: E402,F401
