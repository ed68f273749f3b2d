"""MNIST CNNs from the EventGraD paper.

CNN-2 (live model of dmnist/event): /root/reference/dmnist/event/event.cpp:51-83
  Conv(1->10,k3) -> maxpool2 -> ReLU -> Conv(10->20,k3) -> Dropout2d -> maxpool2 -> ReLU
  -> view(-1,500) -> Linear(500,50) -> ReLU -> dropout(0.5) -> Linear(50,10) -> log_softmax
  8 tensors, 27 480 elements.
CNN-1 (commented out in the reference, kept selectable): event.cpp:15-48
  Conv(1->10,k5), Conv(10->20,k5), Linear(320,100), Linear(100,10): 38 390 elements.
Both return log-probabilities; the training loop applies log_softmax again
(event.cpp:291), which is idempotent.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..ops.shadow import ShadowConv2d, ShadowLinear


class _MnistCNN(nn.Module):
    def __init__(self, k: int, flat: int, hidden: int):
        super().__init__()
        self.conv1 = ShadowConv2d(1, 10, k)
        self.conv2 = ShadowConv2d(10, 20, k)
        self.conv2_drop = nn.Dropout2d()
        self.fc1 = ShadowLinear(flat, hidden)
        self.fc2 = ShadowLinear(hidden, 10)
        self._flat = flat

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = F.relu(F.max_pool2d(self.conv1(x), 2))
        x = F.relu(F.max_pool2d(self.conv2_drop(self.conv2(x)), 2))
        x = x.reshape(-1, self._flat)
        x = F.relu(self.fc1(x))
        x = F.dropout(x, 0.5, self.training)
        x = self.fc2(x)
        return F.log_softmax(x, dim=1)


class CNN2(_MnistCNN):
    def __init__(self):
        super().__init__(k=3, flat=500, hidden=50)


class CNN1(_MnistCNN):
    def __init__(self):
        super().__init__(k=5, flat=320, hidden=100)
