"""Point-wise conv stacks with the reference's parameter names
(core/unopose/model/pointnet2/pytorch_utils.py:25-48, 84-260):
    SharedMLP:  layer{i}.conv.weight, layer{i}.normlayer.bn.{weight,bias,running_mean,running_var,...}
    Conv1d:     conv.weight, conv.bias
so released checkpoints load unchanged."""
import torch.nn as nn


class _NormWrap(nn.Sequential):
    """BatchNorm held under the child name `bn` (pytorch_utils.py:51-81)."""

    def __init__(self, channels, dims):
        super().__init__()
        bn = (nn.BatchNorm1d if dims == 1 else nn.BatchNorm2d)(channels)
        nn.init.constant_(bn.weight, 1.0)
        nn.init.constant_(bn.bias, 0.0)
        self.add_module("bn", bn)


class PointwiseConv(nn.Sequential):
    """1x1 conv -> [BN] -> [activation] with children `conv`, `normlayer`, `activation`."""

    def __init__(self, in_size, out_size, dims=2, bn=False, activation=True):
        super().__init__()
        conv_cls = nn.Conv1d if dims == 1 else nn.Conv2d
        conv = conv_cls(in_size, out_size, kernel_size=1, bias=not bn)
        nn.init.kaiming_normal_(conv.weight)
        if conv.bias is not None:
            nn.init.constant_(conv.bias, 0.0)
        self.add_module("conv", conv)
        if bn:
            self.add_module("normlayer", _NormWrap(out_size, dims))
        if activation:
            self.add_module("activation", nn.ReLU(inplace=True))


class SharedMLP(nn.Sequential):
    """Stack of 1x1 Conv2d(+BN+ReLU) named layer0, layer1, ... (pytorch_utils.py:25-48)."""

    def __init__(self, channels, bn=False):
        super().__init__()
        for i in range(len(channels) - 1):
            self.add_module("layer%d" % i, PointwiseConv(channels[i], channels[i + 1], dims=2, bn=bn, activation=True))


class Conv1d(PointwiseConv):
    """pytorch_utils.py:138-174 restricted to what PositionalEncoding uses (kernel 1)."""

    def __init__(self, in_size, out_size, activation=True, bn=False):
        super().__init__(in_size, out_size, dims=1, bn=bool(bn), activation=bool(activation))
