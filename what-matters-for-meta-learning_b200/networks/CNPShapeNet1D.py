"""Drop-in for the reference's ``networks/CNPShapeNet1D.py``: CNP for ShapeNet1D (networks/CNPShapeNet1D.py:24-143)."""
from networks._families import ShapeNet1DFamilyNP


class CNPShapeNet1D(ShapeNet1DFamilyNP):
    def __init__(self, config):
        super().__init__(config, False)
