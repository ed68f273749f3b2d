"""LeNet-style CIFAR model (dead code in the reference, selectable here).

Reference: /root/reference/dcifar10/common/nnet.hpp:3-33 -- Conv(3->6,k5), Conv(6->16,k5),
Dropout2d, Linear(400,120), Linear(120,84), Linear(84,10), log_softmax.
10 tensors, 62 006 elements. Note the op order differs from the MNIST CNNs:
relu happens before max_pool here.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..ops.shadow import ShadowConv2d, ShadowLinear


class LeNetCifar(nn.Module):
    def __init__(self, classes: int = 10):
        super().__init__()
        self.conv1 = ShadowConv2d(3, 6, 5)
        self.conv2 = ShadowConv2d(6, 16, 5)
        self.conv2_drop = nn.Dropout2d()
        self.fc1 = ShadowLinear(16 * 5 * 5, 120)
        self.fc2 = ShadowLinear(120, 84)
        self.fc3 = ShadowLinear(84, classes)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = F.max_pool2d(F.relu(self.conv1(x)), 2)
        x = F.max_pool2d(F.relu(self.conv2_drop(self.conv2(x))), 2)
        x = x.reshape(-1, 16 * 5 * 5)
        x = F.relu(self.fc1(x))
        x = F.relu(self.fc2(x))
        return F.log_softmax(self.fc3(x), dim=1)
