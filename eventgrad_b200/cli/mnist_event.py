"""Entrypoint equivalent to /root/reference/dmnist/event/event.cpp (EventGraD, CNN-2, MNIST).

Launch: torchrun --nproc-per-node R -m eventgrad_b200.cli.mnist_event [reference positional args] [flags]
(replaces `mpirun -np R ./...`).  See eventgrad_b200/config.py for the CLI contract.
"""
from ._main import run


def main(argv=None):
    return run("mnist_event", argv)


if __name__ == "__main__":
    main()
