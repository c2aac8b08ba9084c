"""Drop-in for the reference's ``networks/ANPShapeNet1D.py``: ANP for ShapeNet1D (networks/ANPShapeNet1D.py:24-161)."""
from networks._families import ShapeNet1DFamilyNP


class ANPShapeNet1D(ShapeNet1DFamilyNP):
    def __init__(self, config):
        super().__init__(config, True)
