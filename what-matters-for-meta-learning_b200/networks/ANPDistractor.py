"""Drop-in for the reference's ``networks/ANPDistractor.py``: ANP for Distractor (networks/ANPDistractor.py:26-135)."""
from networks._families import ResNetFamilyNP


class ANPDistractor(ResNetFamilyNP):
    def __init__(self, config):
        super().__init__(config, True, True)
