"""Drop-in for the reference's ``networks/CNPDistractor.py``: CNP for Distractor (networks/CNPDistractor.py:23-124)."""
from networks._families import ResNetFamilyNP


class CNPDistractor(ResNetFamilyNP):
    def __init__(self, config):
        super().__init__(config, True, False)
