"""DeepONetModules.py surface: FFN trunk and DeepOnetNoBiasOrg (DeepONetModules.py:128-185)."""
from __future__ import annotations

import torch
import torch.nn as nn


def kaiming_init(m):
    if type(m) == nn.Linear:
        torch.nn.init.kaiming_uniform_(m.weight.data, a=0.01, nonlinearity="leaky_relu")
        torch.nn.init.zeros_(m.bias.data)


_ACTIVATIONS = {
    "tanh": nn.Tanh, "Tanh": nn.Tanh,
    "relu": lambda: nn.ReLU(inplace=True), "ReLU": lambda: nn.ReLU(inplace=True),
    "leaky_relu": lambda: nn.LeakyReLU(inplace=True),
    "sigmoid": nn.Sigmoid, "Sigmoid": nn.Sigmoid,
    "softplus": lambda: nn.Softplus(beta=4), "Softplus": lambda: nn.Softplus(beta=4),
    "celu": nn.CELU, "CeLU": nn.CELU, "elu": nn.ELU, "mish": nn.Mish,
}


def activation(name):
    if name not in _ACTIVATIONS:
        raise ValueError("Unknown activation function")
    return _ACTIVATIONS[name]()


class FFN(nn.Module):
    """Trunk net on the fixed grid: act(in) -> (n_hidden-1) x BN1d(act(dropout(Linear))) -> Linear.
    BatchNorm comes AFTER the activation; LeakyReLU slope 0.01 (Q11)."""

    def __init__(self, input_dimension, output_dimension, n_hidden_layers, neurons, act_string, dropout_rate):
        super().__init__()
        self.input_dimension, self.output_dimension = input_dimension, output_dimension
        self.n_hidden_layers, self.neurons = n_hidden_layers, neurons
        self.act_string, self.dropout_rate = act_string, dropout_rate
        self.input_layer = nn.Linear(input_dimension, neurons)
        self.hidden_layers = nn.ModuleList([nn.Linear(neurons, neurons) for _ in range(n_hidden_layers - 1)])
        self.batch_layers = nn.ModuleList([nn.BatchNorm1d(neurons) for _ in range(n_hidden_layers - 1)])
        self.output_layer = nn.Linear(neurons, output_dimension)
        self.activation = activation(act_string)
        self.dropout = nn.Dropout(dropout_rate)
        self.apply(kaiming_init)

    def forward(self, x):
        x = self.activation(self.input_layer(x))
        for lin, bn in zip(self.hidden_layers, self.batch_layers):
            x = bn(self.activation(self.dropout(lin(x))))
        return self.output_layer(x)


class DeepOnetNoBiasOrg(nn.Module):
    """(branch(u) @ trunk(x).T + b0) / sqrt(p)."""

    def __init__(self, branch, trunk):
        super().__init__()
        self.branch, self.trunk = branch, trunk
        self.b0 = nn.Parameter(torch.tensor(0.0), requires_grad=True)
        self.p = self.trunk.output_dimension

    def forward(self, u_, x_):
        weights = self.branch(u_)
        basis = self.trunk(x_)
        return (torch.matmul(weights, basis.T) + self.b0) / self.p ** 0.5
