"""Drop-in for the reference's ``networks/ANP.py``: ANP for ShapeNet3D pose regression (networks/ANP.py:25-130)."""
from networks._families import ResNetFamilyNP


class ANP(ResNetFamilyNP):
    def __init__(self, config):
        super().__init__(config, False, True)
