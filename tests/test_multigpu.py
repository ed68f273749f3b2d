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


def _run(world, *args, timeout=300, env=None):
    _PORT[0] += 1
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_PORT[0]),
           os.path.join(ROOT, "tests", "dist_worker.py"), *args]
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=e)
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


def test_decent_double_buffered_is_the_default_and_single_slot_still_works():
    """csrc/gossip_dbuf.cu (two inbox slots, no WAR ack) over real NVLink peers, and the single-slot + ack protocol."""
    for w in _worlds():
        assert "dbuf=1" in _run(w, "--algo", "decent", "--backend", "p2p", "--steps", "11")
        assert "dbuf=0" in _run(w, "--algo", "decent", "--backend", "p2p", "--steps", "11", "--no-double-buffer")


def test_ce_push_split_step():
    """Copy-engine push (csrc/ce_push.cu) as the first half of the split decent step."""
    for w in _worlds():
        assert "ce_push=1" in _run(w, "--algo", "decent", "--backend", "p2p", "--steps", "10", "--overlap", "--ce-push")


@pytest.mark.parametrize("algo", ["cent", "decent"])
def test_nvls_allreduce(algo):
    """EGB_NVLS=1: window in torch symmetric memory, csrc/allreduce_nvls.cu (multimem.ld_reduce / multimem.st through
    the NVSwitch) for the cent step on the ResNet arena (two-shot) and for the final parameter averaging."""
    ws = _worlds()
    if not ws:
        pytest.skip("needs >= 2 GPUs")
    out = _run(ws[-1], "--algo", algo, "--backend", "p2p", "--model", "resnet18", "--dataset", "cifar10", "--steps", "3",
               env={"EGB_NVLS": "1"}, timeout=600)
    if "nvls=1" not in out:
        pytest.skip("no multicast support on this fabric (window fell back to plain peer mappings)")
    if algo == "cent":
        assert "nvls_step=1" in out
