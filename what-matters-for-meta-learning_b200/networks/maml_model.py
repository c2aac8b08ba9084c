"""``Model`` base of the MMAML networks: an ``nn.Module`` exposing its parameters as an ordered name -> tensor mapping
(``param_dict``), which the inner loop rewrites functionally (reference networks/maml_model.py)."""
from collections import OrderedDict

import torch


class Model(torch.nn.Module):
    @property
    def param_dict(self):
        return OrderedDict(self.named_parameters())
