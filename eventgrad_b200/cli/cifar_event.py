"""Entrypoint equivalent to /root/reference/dcifar10/event/event.cpp (EventGraD, ResNet, CIFAR-10).

Launch: torchrun --nproc-per-node R -m eventgrad_b200.cli.cifar_event [reference positional args] [flags]
(replaces `mpirun -np R ./...`).  See eventgrad_b200/config.py for the CLI contract.
"""
from ._main import run


def main(argv=None):
    return run("cifar_event", argv)


if __name__ == "__main__":
    main()
