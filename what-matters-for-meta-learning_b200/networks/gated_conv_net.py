"""Drop-in ``networks.gated_conv_net``.

``B200NP_MMAML=1`` selects the B200 ``GatedConvModel`` below: same constructor, parameter / buffer names
(``features.layer{i}_conv.*``, ``features.layer{i}_bn.running_*``, ``classifier.fully_connected.*``), initialisation
(xavier-uniform weights, zero biases, networks/gated_conv_net.py:15-19) and ``forward(x, params=None, embeddings=None)``
contract as the reference, with the arithmetic in libb200np (b200np/mmaml.py).  Second order is supported: under
``torch.autograd.grad(..., create_graph=True)`` -- what ``MetaLearner(first_order=False)`` does (train.py:99,
trainer/meta_learner_reg.py:116-122) -- the gradients come back with a graph built from the differentiable backward ops
of b200np/second_order.py, so the outer gradient contains the second-order terms (tests/test_mmaml.py runs the
reference's meta-step against golden vectors of the reference).  Without the variable -- the default -- this module
hands out the reference's own classes.
"""
import os
from collections import OrderedDict

import torch
import torch.nn.functional as F

from ._refload import reference_module
from .maml_model import Model


def weight_init(module):
    if isinstance(module, (torch.nn.Linear, torch.nn.Conv2d)):
        torch.nn.init.xavier_uniform_(module.weight)
        module.bias.data.zero_()


class GatedConvModel(Model):
    def __init__(self, input_channels, output_size, num_channels=64, kernel_size=3, padding=1, nonlinearity=F.relu,
                 use_max_pool=False, img_side_len=28, condition_type='affine', condition_order='low2high',
                 verbose=False):
        super().__init__()
        if use_max_pool or kernel_size != 3 or padding != 1 or condition_type != 'affine':
            raise NotImplementedError("B200 GatedConvModel covers the configuration the reference uses "
                                      "(networks/MMAMLShapeNet1D.py:47-55): stride-2 3x3 convs, affine FiLM, no max-pool")
        self._input_channels, self._output_size, self._num_channels = input_channels, output_size, num_channels
        self._kernel_size, self._nonlinearity, self._use_max_pool = kernel_size, nonlinearity, use_max_pool
        self._padding, self._condition_type, self._condition_order = padding, condition_type, condition_order
        self._bn_affine, self._reuse, self._verbose = False, False, verbose
        self._conv_stride = 2
        self._features_size = (img_side_len // 14) ** 2
        layers, cin = [], input_channels
        for i in range(1, 5):
            cout = num_channels * 2 ** (i - 1)
            layers += [(f'layer{i}_conv', torch.nn.Conv2d(cin, cout, kernel_size, stride=2, padding=padding)),
                       (f'layer{i}_bn', torch.nn.BatchNorm2d(cout, affine=False, momentum=0.001)),
                       (f'layer{i}_condition', torch.nn.ReLU(inplace=True)),
                       (f'layer{i}_relu', torch.nn.ReLU(inplace=True))]
            cin = cout
        self.features = torch.nn.Sequential(OrderedDict(layers))
        self.classifier = torch.nn.Sequential(OrderedDict([
            ('fully_connected', torch.nn.Linear(num_channels * 8, output_size))]))
        self.apply(weight_init)

    def forward(self, x, params=None, embeddings=None):
        from b200np import mmaml
        if params is None:
            params = OrderedDict(self.named_parameters())
        return mmaml.gated_conv_forward(self, x, params, embeddings)


if os.environ.get("B200NP_MMAML", "0") != "1":
    _ref = reference_module("gated_conv_net")
    if _ref is not None:
        GatedConvModel = _ref.GatedConvModel          # noqa: F811  (default: the reference's own class)
        weight_init = _ref.weight_init                # noqa: F811
