"""Opt-in checks for code that has not been run on hardware yet (EGB_EXPERIMENTAL=1 to enable).
They are skipped in the default GPU tier so an unvalidated kernel can never turn the suite red."""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("EGB_EXPERIMENTAL") != "1", reason="experimental: set EGB_EXPERIMENTAL=1")]


@pytest.mark.parametrize("M,K,N", [(128, 64, 128), (1000, 784, 128), (300, 128, 256), (77, 16, 16), (60000, 784, 128)])
@pytest.mark.parametrize("relu", [False, True])
def test_linear_tc_tma_variant(M, K, N, relu, monkeypatch):
    """csrc/linear_tc_tma.cu (TMA + SWIZZLE_128B + persistent CTAs) against the fp32 reference."""
    from eventgrad_b200.ops.linear_tc import linear_tc_forward
    monkeypatch.setenv("EGB_TC_LINEAR", "tma")
    g = torch.Generator(device="cuda").manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g, device="cuda") / K ** 0.5).to(torch.bfloat16)
    b = torch.randn(N, generator=g, device="cuda")
    y = linear_tc_forward(x, w, b, relu, torch.float32)
    torch.cuda.synchronize()
    ref = x.float() @ w.float().t() + b
    if relu:
        ref = ref.relu()
    torch.testing.assert_close(y, ref, rtol=2e-3, atol=2e-3)


def test_native_loader_on_cuda_path():
    """--native-loader on: C++ prefetch thread feeding pinned slots + async H2D (CPU-validated only so far)."""
    from eventgrad_b200.data import BatchLoader, ShardSampler, synthetic_source
    src = synthetic_source("cifar10", 1000).pin()
    a = BatchLoader(src, ShardSampler(1000, 1, 0, "sequential"), 64, "cuda", native="on")
    b = BatchLoader(src, ShardSampler(1000, 1, 0, "sequential"), 64, "cuda", native="off")
    assert a.native is not None
    for (xa, ya), (xb, yb) in zip(a, b):
        assert torch.equal(xa, xb) and torch.equal(ya, yb)


def test_p2p_file_write_logs(tmp_path):
    """Reference debug files from the device log ring of the p2p backend (MNIST self-loop, 1 GPU)."""
    from eventgrad_b200.config import preset
    from eventgrad_b200.data import synthetic_source
    from eventgrad_b200.engine.trainer import Trainer
    from eventgrad_b200.utils.dist import DistEnv
    cfg = preset("mnist_event", backend="p2p", device="cuda", train_samples=640, test_samples=128, epochs=1,
                 quiet=True, file_write=1, log_dir=str(tmp_path))
    tr = Trainer(cfg, DistEnv(0, 1, 0, torch.device("cuda", 0), "none"), train_source=synthetic_source("mnist", 640).pin(),
                 test_source=synthetic_source("mnist", 128, train=False).pin())
    tr.fit()
    tr.finalize()
    send = open(tmp_path / "send0.txt").read().splitlines()
    recv = open(tmp_path / "recv0.txt").read().splitlines()
    assert len(send) == 10 and len(recv) == 10 and len(send[0].split(",  ")) == 8 * 3 + 1
    tr.close()


@pytest.fixture
def bn_v2(monkeypatch):
    """Route FusedBNAct through csrc/bn_act_v2.cu (ReLU bit mask) for one test."""
    from eventgrad_b200.ops import bn_act
    monkeypatch.setenv("EGB_BN_V2", "1")
    bn_act._WS.clear()                      # the switch is read when the per-device workspace is created
    yield
    bn_act._WS.clear()


@pytest.mark.parametrize("shape", [(32, 64, 32, 32), (7, 128, 16, 16), (5, 256, 8, 8), (3, 512, 4, 4), (2, 2048, 4, 4),
                                   (1, 64, 3, 5), (128, 64, 32, 32), (64, 128, 16, 16), (256, 64, 32, 32)])
@pytest.mark.parametrize("relu,res", [(True, False), (True, True), (False, False), (False, True)])
def test_bn_v2_forward_backward(shape, relu, res, bn_v2):
    """Same acceptance test as the default kernels (tests/test_gpu_bn.py) on the v2 code path."""
    from test_gpu_bn import test_fused_bn_act_forward_backward as body
    from eventgrad_b200.ops import bn_act
    body(shape, relu, res)
    assert bn_act._WS[torch.device("cuda", torch.cuda.current_device())]["v2"] is True


def test_bn_v2_matches_v1_bitwise_forward(bn_v2):
    """v2's forward is the same arithmetic as the split v1 path: y must be bit-identical; dx within bf16."""
    from eventgrad_b200.ops import bn_act
    from eventgrad_b200.ops.bn_act import FusedBNAct
    from test_gpu_bn import _mk
    x, r, dy = _mk(128, 128, 16, 16, seed=5)
    outs = []
    for v2 in (True, False):
        os.environ["EGB_BN_V2"] = "1" if v2 else "0"
        bn_act._WS.clear()
        bn = FusedBNAct(128).cuda().train()
        xa, ra = x.clone().requires_grad_(True), r.clone().requires_grad_(True)
        y = bn(xa, residual=ra, relu=True)
        y.backward(dy)
        outs.append((y.detach().clone(), xa.grad.clone(), ra.grad.clone(), bn.weight.grad.clone(), bn.bias.grad.clone()))
    a, b = outs
    assert torch.equal(a[0], b[0])
    assert torch.equal(a[2], b[2])                                   # dres = masked dy: exact in both
    torch.testing.assert_close(a[4], b[4], rtol=1e-5, atol=1e-5)     # dbeta: same sums, different split
    torch.testing.assert_close(a[3], b[3], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(a[1].float(), b[1].float(), rtol=2e-2, atol=2e-3)


def test_bn_v2_resnet_step_and_graph(bn_v2):
    """Flagship step (bf16 NHWC, whole-step CUDA graph, grad table) on the v2 BN kernels: loss finite and
    parameters close to the v1 run from the same seed."""
    from eventgrad_b200.ops import bn_act
    from eventgrad_b200.config import preset
    from eventgrad_b200.data import synthetic_source
    from eventgrad_b200.engine.trainer import Trainer
    from eventgrad_b200.utils.dist import DistEnv
    res = {}
    for v2 in (True, False):
        os.environ["EGB_BN_V2"] = "1" if v2 else "0"
        bn_act._WS.clear()
        cfg = preset("cifar_event", backend="p2p", device="cuda", train_samples=512, test_samples=128, batch_size=64,
                     epochs=100, quiet=True, max_steps=6, augment=False, dtype="bf16", channels_last=True,
                     cuda_graph=True)
        torch.manual_seed(0)
        tr = Trainer(cfg, DistEnv(0, 1, 0, torch.device("cuda", 0), "none"),
                     train_source=synthetic_source("cifar10", 512).pin(),
                     test_source=synthetic_source("cifar10", 128, train=False).pin())
        tr.fit()
        tr.backend.check_status()
        res[v2] = (float(tr.last_loss), tr.arena.theta.clone())
        tr.close()
    assert all(l == l and abs(l) < 1e4 for l, _ in res.values())
    assert abs(res[True][0] - res[False][0]) < 0.2 * max(1.0, abs(res[False][0]))
    rel = float((res[True][1] - res[False][1]).norm() / res[False][1].norm())
    assert rel < 2e-3, rel


@pytest.mark.parametrize("R", [1, 2, 3, 4])
@pytest.mark.parametrize("mu", [0.0, 0.9])
@pytest.mark.parametrize("model", ["cnn2", "mlp"])
def test_double_buffered_decent_bitwise_vs_simulator(R, mu, model):
    """csrc/gossip_dbuf.cu (two inbox slots, no WAR ack) on R virtual ranks of one GPU: same bits as the
    oracle after 7 steps (odd and even slots both used several times)."""
    from test_gpu_kernels import _cfg, _grads, _mask_pad, _world
    from eventgrad_b200.engine.simulator import RingSimulator
    cfg = _cfg("decent", momentum=mu, model=model, double_buffer=True)
    w = _world(cfg, R, model=model)
    assert all(be.dbuf for be in w.backends)
    t = w.arenas[0].table
    sim = RingSimulator(R, w.arenas[0].theta.cpu(), t, "decent", lr=cfg.lr, momentum=mu, serial_skip=False)
    for s in range(7):
        g = _mask_pad(w, _grads(R, t.n_padded, 300 + s))
        w.step(g)
        sim.step([x.cpu() for x in g])
    torch.cuda.synchronize()
    for be in w.backends:
        be.check_status()
    for r in range(R):
        assert torch.equal(w.arenas[r].theta.cpu(), sim.theta[r]), f"rank {r} theta differs"
    w.close()


_PORT = [29880]


def _torchrun(world, *args, env=None, timeout=600):
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    _PORT[0] += 1
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_PORT[0]),
           os.path.join(root, "tests", "dist_worker.py"), *args]
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=root, env=e)
    out = r.stdout + r.stderr
    assert r.returncode == 0 and "WORKER_OK" in out, out[-3000:]
    return out


@pytest.mark.multigpu
def test_double_buffered_decent_multi_gpu():
    """csrc/gossip_dbuf.cu over real NVLink peers: bit-exact vs the simulator."""
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    for w in [x for x in (2, 4, 8) if x <= n]:
        out = _torchrun(w, "--algo", "decent", "--backend", "p2p", "--steps", "11", "--double-buffer")
        assert "dbuf=1" in out


@pytest.mark.multigpu
@pytest.mark.parametrize("algo", ["cent", "decent"])
def test_nvls_allreduce_multi_gpu(algo):
    """EGB_NVLS=1: window in torch symmetric memory, csrc/allreduce_nvls.cu (multimem.ld_reduce/st) for the
    cent step on the ResNet arena (two-shot) and for the final parameter averaging."""
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    w = max(x for x in (2, 4, 8) if x <= n)
    out = _torchrun(w, "--algo", algo, "--backend", "p2p", "--model", "resnet18", "--dataset", "cifar10", "--steps", "3",
                    env={"EGB_NVLS": "1"})
    if "nvls=1" not in out:
        pytest.skip("no multicast support on this fabric (window fell back to plain peer mappings)")
    if algo == "cent":
        assert "nvls_step=1" in out


def test_conv_split_backward_matches_default(monkeypatch):
    """ops/shadow.py:_ConvSplitBwd (wgrad on a side stream, dgrad on the main stream) vs the default autograd
    path of the same ShadowConv2d: same cuDNN kernels -> gradients equal to bf16 precision, run 20 times to give
    a stream-ordering bug a chance to show."""
    from eventgrad_b200.ops.shadow import ShadowConv2d
    torch.manual_seed(1)
    conv = ShadowConv2d(64, 128, 3, stride=2, padding=1, bias=False).cuda()
    conv.w16 = conv.weight.detach().to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    x0 = torch.randn(32, 64, 16, 16, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    dy = None
    res = {}
    for split in ("0", "1"):
        monkeypatch.setenv("EGB_CONV_SPLIT_BWD", split)
        outs = []
        for _ in range(20 if split == "1" else 1):
            x = x0.clone().requires_grad_(True)
            conv.w16.grad = None
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = conv(x)
            if dy is None:
                dy = torch.randn_like(y)
            y.backward(dy)
            torch.cuda.synchronize()
            outs.append((y.detach().float(), x.grad.float().clone(), conv.w16.grad.float().clone()))
        res[split] = outs
    ref = res["0"][0]
    for y, dx, dw in res["1"]:
        torch.testing.assert_close(y, ref[0], rtol=1e-2, atol=1e-2)
        torch.testing.assert_close(dx, ref[1], rtol=2e-2, atol=2e-2)
        torch.testing.assert_close(dw, ref[2], rtol=2e-2, atol=5e-2)


def test_conv_split_backward_resnet_graph_step(monkeypatch):
    """Whole flagship step (per-GPU batch 32, CUDA graph, grad table) with the split backward: parameters after
    6 steps track the default run from the same seed."""
    from eventgrad_b200.config import preset
    from eventgrad_b200.data import synthetic_source
    from eventgrad_b200.engine.trainer import Trainer
    from eventgrad_b200.utils.dist import DistEnv
    res = {}
    for split in ("1", "0"):
        monkeypatch.setenv("EGB_CONV_SPLIT_BWD", split)
        cfg = preset("cifar_event", backend="p2p", device="cuda", train_samples=256, test_samples=64, batch_size=32,
                     epochs=100, quiet=True, max_steps=6, augment=False, dtype="bf16", channels_last=True,
                     cuda_graph=True)
        torch.manual_seed(0)
        tr = Trainer(cfg, DistEnv(0, 1, 0, torch.device("cuda", 0), "none"),
                     train_source=synthetic_source("cifar10", 256).pin(),
                     test_source=synthetic_source("cifar10", 64, train=False).pin())
        tr.fit()
        tr.backend.check_status()
        res[split] = (float(tr.last_loss), tr.arena.theta.clone())
        tr.close()
    assert all(l == l and abs(l) < 1e4 for l, _ in res.values())
    rel = float((res["1"][1] - res["0"][1]).norm() / res["0"][1].norm())
    assert rel < 2e-3, rel


@pytest.mark.parametrize("R", [1, 2, 3])
def test_ce_push_split_step_matches_simulator(R):
    """csrc/ce_push.cu: ack wait kernel -> 2 x cudaMemcpyAsync -> pushed-flag kernel as the push half of the split
    step (decent), R virtual ranks on one GPU, bit-exact vs the oracle."""
    from test_gpu_kernels import _cfg, _grads, _mask_pad, _world
    from eventgrad_b200.engine.simulator import RingSimulator
    cfg = _cfg("decent", overlap_push=True, ce_push=True)
    w = _world(cfg, R)
    assert all(be.ce_push for be in w.backends)
    t = w.arenas[0].table
    sim = RingSimulator(R, w.arenas[0].theta.cpu(), t, "decent", lr=cfg.lr, momentum=cfg.momentum, serial_skip=False)
    for s in range(8):
        g = _mask_pad(w, _grads(R, t.n_padded, 700 + s))
        w.step(g)
        sim.step([x.cpu() for x in g])
    torch.cuda.synchronize()
    for r, be in enumerate(w.backends):
        be.check_status()
        assert torch.equal(w.arenas[r].theta.cpu(), sim.theta[r]), f"rank {r}"
    w.close()


@pytest.mark.multigpu
def test_ce_push_multi_gpu():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    for w in [x for x in (2, 4, 8) if x <= n]:
        out = _torchrun(w, "--algo", "decent", "--backend", "p2p", "--steps", "10", "--overlap", "--ce-push")
        assert "ce_push=1" in out


@pytest.fixture
def bn_cluster(monkeypatch):
    """FusedBNAct through csrc/bn_act_cluster.cu (single launch, thread-block cluster + DSMEM) where the slice fits."""
    from eventgrad_b200.ops import bn_act
    monkeypatch.setenv("EGB_BN_V2", "1")
    monkeypatch.setenv("EGB_BN_CLUSTER", "1")
    bn_act._WS.clear()
    yield
    bn_act._WS.clear()


# (N, C, H, W): M = N*H*W rows.  cluster sizes exercised: 1 (M<=256), 2, 4, 8, 16 (M=16384 fwd: 8; bwd: 16 or fallback)
@pytest.mark.parametrize("shape", [(1, 64, 3, 5), (2, 2048, 4, 4), (3, 512, 4, 4), (32, 512, 4, 4), (32, 256, 8, 8),
                                   (32, 128, 16, 16), (7, 128, 16, 16), (64, 128, 16, 16), (32, 64, 32, 32)])
@pytest.mark.parametrize("relu,res", [(True, False), (True, True), (False, False), (False, True)])
def test_bn_cluster_forward_backward(shape, relu, res, bn_cluster):
    """Same acceptance test as the default kernels on the cluster path; the last shape (32768 rows) must fall back."""
    from test_gpu_bn import test_fused_bn_act_forward_backward as body
    from eventgrad_b200.ops import bn_act
    body(shape, relu, res)
    ws = bn_act._WS[torch.device("cuda", torch.cuda.current_device())]
    M = shape[0] * shape[2] * shape[3]
    if M <= 8192:
        assert ws["cluster_taken"] >= 1, "cluster kernel was not used for a shape that fits"
    if M > 16 * 1536:
        assert ws["cluster_taken"] == 0


def test_bn_cluster_resnet_step_b32(bn_cluster):
    """Flagship step at the per-GPU batch of the 8-GPU configuration with cluster BN: tracks the default run."""
    from eventgrad_b200.ops import bn_act
    from eventgrad_b200.config import preset
    from eventgrad_b200.data import synthetic_source
    from eventgrad_b200.engine.trainer import Trainer
    from eventgrad_b200.utils.dist import DistEnv
    res = {}
    for mode in ("cluster", "default"):
        os.environ["EGB_BN_V2"] = "1" if mode == "cluster" else "0"
        os.environ["EGB_BN_CLUSTER"] = "1" if mode == "cluster" else "0"
        bn_act._WS.clear()
        cfg = preset("cifar_event", backend="p2p", device="cuda", train_samples=256, test_samples=64, batch_size=32,
                     epochs=100, quiet=True, max_steps=6, augment=False, dtype="bf16", channels_last=True,
                     cuda_graph=True)
        torch.manual_seed(0)
        tr = Trainer(cfg, DistEnv(0, 1, 0, torch.device("cuda", 0), "none"),
                     train_source=synthetic_source("cifar10", 256).pin(),
                     test_source=synthetic_source("cifar10", 64, train=False).pin())
        tr.fit()
        tr.backend.check_status()
        if mode == "cluster":
            assert bn_act._WS[torch.device("cuda", 0)]["cluster_taken"] > 0
        res[mode] = (float(tr.last_loss), tr.arena.theta.clone())
        tr.close()
    assert all(l == l and abs(l) < 1e4 for l, _ in res.values())
    rel = float((res["cluster"][1] - res["default"][1]).norm() / res["default"][1].norm())
    assert rel < 2e-3, rel
