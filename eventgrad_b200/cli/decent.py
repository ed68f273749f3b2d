"""Entrypoint equivalent to /root/reference/dmnist/decent/decent.cpp (D-PSGD ring gossip, MLP, MNIST).

Launch: torchrun --nproc-per-node R -m eventgrad_b200.cli.decent [reference positional args] [flags]
(replaces `mpirun -np R ./...`).  See eventgrad_b200/config.py for the CLI contract.
"""
from ._main import run


def main(argv=None):
    return run("decent", argv)


if __name__ == "__main__":
    main()
