"""Model registry. Models are plain PyTorch modules (the north-star keeps the model /
autograd layer in PyTorch); what is native is everything that happens to their
parameters between backward() and the next forward()."""
from __future__ import annotations

import torch.nn as nn

from .cnn import CNN1, CNN2
from .lenet import LeNetCifar
from .mlp import MLP
from .resnet import BasicBlock, BottleNeck, ResNet, make_resnet, _SPECS

MODEL_NAMES = ("mlp", "cnn1", "cnn2", "lenet") + tuple(_SPECS)


def build_model(name: str, *, resnet_variant: str = "ref", classes: int = 10) -> nn.Module:
    name = name.lower()
    if name == "mlp":
        return MLP(classes=classes)
    if name == "cnn1":
        return CNN1()
    if name == "cnn2":
        return CNN2()
    if name == "lenet":
        return LeNetCifar(classes)
    if name in _SPECS:
        return make_resnet(name, classes, resnet_variant)
    raise ValueError(f"unknown model {name!r}; choose from {MODEL_NAMES}")


def outputs_log_probs(name: str) -> bool:
    """CNN-1/2 and LeNet end in log_softmax; MLP and ResNet return logits."""
    return name.lower() in ("cnn1", "cnn2", "lenet")


def param_inventory(model: nn.Module):
    """(n_tensors, n_elements, [(name, numel), ...]) in named_parameters() order."""
    items = [(n, p.numel()) for n, p in model.named_parameters()]
    return len(items), sum(n for _, n in items), items
