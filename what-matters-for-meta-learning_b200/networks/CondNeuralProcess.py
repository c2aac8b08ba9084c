"""Drop-in for the reference's ``networks/CondNeuralProcess.py``: CNP for ShapeNet3D (networks/CondNeuralProcess.py:26-123)."""
from networks._families import ResNetFamilyNP


class CondNeuralProcess(ResNetFamilyNP):
    def __init__(self, config):
        super().__init__(config, False, False)
