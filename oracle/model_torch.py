"""CPU port of the reference's model compositions -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.

Mirrors examples/pytorch_based/pytorch_hcp_tgcn.py:93-155 and pytorch_mnist_tgcn.py:67-92 with the
conv layers and pooling performed by oracle/layers_torch.py (same ATen calls as the reference).
Parameter names match tgcn_b200.workloads models so a state_dict can be copied across.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import layers_torch as ref


def _uniform(t, size):
    t.data.uniform_(-1.0 / math.sqrt(size), 1.0 / math.sqrt(size))


class _Conv(nn.Module):
    def __init__(self, L, wshape, bshape, fan):
        super().__init__()
        self.L = L
        self.weight = nn.Parameter(torch.empty(*wshape))
        self.bias = nn.Parameter(torch.empty(*bshape))
        _uniform(self.weight, fan)
        _uniform(self.bias, fan)

    def forward(self, x):
        return ref.cheb_layer(self.L, x, self.weight, self.bias)


class PortNetTGCN_HCP(nn.Module):
    """pytorch_hcp_tgcn.py:93-155, call for call: real part of the FFT along the time window (:133, written with the
    removed `torch.rfft`; `torch.fft.fft(x, dim=2).real` is its replacement), tgcn1, relu, drop1 (0.1), pool4, gcn2,
    relu, pool4, view, fc1, dense1_bn, relu, drop2 (0.5), fc2, log_softmax."""

    def __init__(self, L, horizon=15, K=10, g1=32, g2=64, hidden=200, n_classes=6, time_dft=True, drop1=0.1, drop2=0.5):
        super().__init__()
        self.tgcn1 = _Conv(L[0], (K, horizon, 1, g1), (1, L[0].shape[0], g1), K)
        self.drop1 = nn.Dropout(drop1)
        self.gcn2 = _Conv(L[2], (K, g1, g2), (1, 1, g2), K * g1)
        self.fc1 = nn.Linear(int(L[2].shape[0] * g2 / 4), hidden)
        self.dense1_bn = nn.BatchNorm1d(hidden)
        self.drop2 = nn.Dropout(drop2)
        self.fc2 = nn.Linear(hidden, n_classes)
        self.time_dft = time_dft

    def forward(self, x):
        if self.time_dft:
            x = torch.fft.fft(x, dim=2).real
        x = ref.pool(self.drop1(F.relu(self.tgcn1(x))), 4)
        x = ref.pool(F.relu(self.gcn2(x)), 4)
        x = x.reshape(x.shape[0], -1)
        x = self.drop2(F.relu(self.dense1_bn(self.fc1(x))))
        return F.log_softmax(self.fc2(x), dim=1)


class PortNetTGCN_MNIST(nn.Module):
    def __init__(self, L, horizon=12, K=10, g1=15, n_classes=10):
        super().__init__()
        self.tgcn1 = _Conv(L[0], (K, horizon, 1, g1), (1, L[0].shape[0], g1), K)
        self.fc1 = nn.Linear(L[0].shape[0] * g1, n_classes)

    def forward(self, x):
        x = F.relu(self.tgcn1(x))
        return F.log_softmax(self.fc1(x.reshape(x.shape[0], -1)), dim=1)
