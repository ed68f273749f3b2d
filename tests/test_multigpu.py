"""Real multi-GPU checks: one process per GPU (torchrun), CUDA-IPC windows over NVLink, fused
kernels vs the simulator (bitwise in iter-sync mode) and the NCCL baseline backend."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_PORT = [29700]


def _run(world, *args, timeout=300):
    _PORT[0] += 1
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_PORT[0]),
           os.path.join(ROOT, "tests", "dist_worker.py"), *args]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    out = r.stdout + r.stderr
    assert r.returncode == 0 and "WORKER_OK" in out, out[-3000:]
    return out


def _worlds():
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    want = os.environ.get("EGB_TEST_WORLDS")
    cand = [int(x) for x in want.split(",")] if want else [2, 4, 8]
    return [w for w in cand if w <= n]


@pytest.mark.parametrize("algo", ["cent", "decent", "event", "spevent"])
def test_p2p_vs_simulator(algo):
    for w in _worlds():
        _run(w, "--algo", algo, "--backend", "p2p", "--steps", "12")


@pytest.mark.parametrize("algo", ["decent", "event", "spevent"])
def test_p2p_overlap_vs_simulator(algo):
    for w in _worlds():
        _run(w, "--algo", algo, "--backend", "p2p", "--steps", "10", "--overlap")


def test_p2p_event_async():
    for w in _worlds()[:1]:
        _run(w, "--algo", "event", "--backend", "p2p", "--sync-mode", "async", "--thres-type", "0",
             "--constant", "0", "--steps", "8")


@pytest.mark.parametrize("algo", ["cent", "decent", "event", "spevent"])
def test_nccl_baseline_vs_simulator(algo):
    for w in _worlds()[:1]:
        _run(w, "--algo", algo, "--backend", "nccl", "--steps", "8")


def test_p2p_resnet_decent():
    for w in _worlds()[:1]:
        _run(w, "--algo", "decent", "--backend", "p2p", "--model", "resnet18", "--dataset", "cifar10",
             "--steps", "3", timeout=600)
