"""Drop-in ``networks.conv_embedding_model``; ``B200NP_MMAML=1`` selects the B200 class, otherwise the
reference's own (see networks/gated_conv_net.py in this package).

``ConvEmbeddingModel``: same constructor signature, parameter names (``conv.conv{i}.*``, ``conv.bn{i}.*``,
``linear.*``, ``_embeddings.{j}.*``) and ``forward(x, params=None, return_task_embedding=False)`` contract as
networks/conv_embedding_model.py:13-191, for the configuration the reference instantiates
(networks/MMAMLShapeNet1D.py:58-75: convolutional, batch-norm, average pool after the convs, no RNN aggregation).
"""
import os
from collections import OrderedDict

import numpy as np  # noqa: F401  (the reference module exposes it too)
import torch

from ._refload import reference_module


class ConvEmbeddingModel(torch.nn.Module):
    def __init__(self, input_size, output_size, embedding_dims, hidden_size=128, num_layers=1, convolutional=False,
                 num_conv=4, num_channels=32, num_channels_max=256, rnn_aggregation=False, linear_before_rnn=False,
                 embedding_pooling='max', batch_norm=True, avgpool_after_conv=True, num_sample_embedding=0,
                 sample_embedding_file='embedding.hdf5', img_size=(1, 28, 28), verbose=False):
        super().__init__()
        if not convolutional or rnn_aggregation or not batch_norm or not avgpool_after_conv or num_sample_embedding:
            raise NotImplementedError("B200 ConvEmbeddingModel covers the configuration of networks/MMAMLShapeNet1D.py:58-75")
        self._input_size, self._output_size, self._hidden_size = input_size, output_size, hidden_size
        self._num_layers, self._embedding_dims, self._bidirectional = num_layers, embedding_dims, True
        self._device, self._convolutional, self._num_conv = 'cpu', convolutional, num_conv
        self._num_channels, self._num_channels_max, self._batch_norm = num_channels, num_channels_max, batch_norm
        self._img_size, self._rnn_aggregation, self._embedding_pooling = img_size, rnn_aggregation, embedding_pooling
        self._linear_before_rnn, self._embeddings_array = linear_before_rnn, []
        self._num_sample_embedding, self._sample_embedding_file = num_sample_embedding, sample_embedding_file
        self._avgpool_after_conv, self._reuse, self._verbose = avgpool_after_conv, False, verbose
        chans = [img_size[0]] + [min(num_channels_max, num_channels * 2 ** i) for i in range(num_conv)]
        chans[0] = min(num_channels_max, chans[0])
        conv_list = OrderedDict()
        for i in range(num_conv):
            conv_list[f'conv{i + 1}'] = torch.nn.Conv2d(chans[i], chans[i + 1], (3, 3), stride=2, padding=1)
            conv_list[f'bn{i + 1}'] = torch.nn.BatchNorm2d(chans[i + 1], momentum=0.001)
            conv_list[f'relu{i + 1}'] = torch.nn.ReLU(inplace=True)
        self.conv = torch.nn.Sequential(conv_list)
        self._num_layer_per_conv = len(conv_list) // num_conv
        self.rnn = None
        self.linear = torch.nn.Linear(chans[-1], hidden_size)
        self.relu_after_linear = torch.nn.ReLU(inplace=True)
        self._embeddings = torch.nn.ModuleList([torch.nn.Linear(hidden_size, dim) for dim in embedding_dims])

    def forward(self, x, params=None, return_task_embedding=False):
        from b200np import mmaml
        if params is None:
            params = OrderedDict(self.named_parameters())
        return mmaml.conv_embedding_forward(self, x, params, return_task_embedding)

    def to(self, device, **kwargs):
        self._device = device
        super().to(device, **kwargs)


if os.environ.get("B200NP_MMAML", "0") != "1":
    _ref = reference_module("conv_embedding_model")
    if _ref is not None:
        ConvEmbeddingModel = _ref.ConvEmbeddingModel  # noqa: F811  (default: the reference's own class)
